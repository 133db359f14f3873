#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02g_pytest.log | cut -c1-300
timeout 120 python scripts/prof_timeline.py --variant full --out $O/r02g_timeline_c2_full.txt > /dev/null 2> $O/r02g_tl.err || tail -3 $O/r02g_tl.err
cut -c1-110 $O/r02g_timeline_c2_full.txt
for env in "X=1" "TIC_CODE_WARM=0" "TIC_PDL_CHAINS=1" "TIC_PDL=1" "TIC_HI_PRIORITY=0"; do
  echo "$env"; env $env timeout 120 python scripts/timeline.py --replays 400 --plain-only --variant full,itc,fusion 2>&1 | grep "^workload"
done
timeout 60 python scripts/fused_in_step.py 2>&1 | grep -A2 "variant itc (" | tail -3
