#!/bin/bash
# Round 2, GPU pass B: where does the c2 step spend its time?  CUPTI kernel timelines + A/B of the restructuring switches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02b_pytest.log
for v in full itc fusion; do
  timeout 120 python scripts/prof_timeline.py --variant $v --out $O/r02b_timeline_c2_$v.txt > /dev/null 2> $O/r02b_tl_$v.err || tail -3 $O/r02b_tl_$v.err
done
cat $O/r02b_timeline_c2_full.txt
echo "--- A/B (plain graph, 400 replays)"
for env in "X=1" "TIC_TIMELINE_SNAPSHOT=1" "TIC_ITC_FUSED_SMALL=0" "TIC_CONCAT_PAIRWISE=0" "TIC_PDL_CHAINS=1" "TIC_TIMELINE_SNAPSHOT=1 TIC_PDL_CHAINS=1" "TIC_HI_PRIORITY=0"; do
  echo "$env"; env $env timeout 120 python scripts/timeline.py --replays 400 --plain-only --variant full,itc,fusion 2>&1 | grep "^workload"
done
