#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02d_pytest.log | cut -c1-300
timeout 300 python scripts/kernel_times.py > $O/r02d_kernel_times_c2.txt 2> $O/r02d_kt.err || tail -5 $O/r02d_kt.err
cat $O/r02d_kernel_times_c2.txt
timeout 120 python scripts/prof_timeline.py --variant full --out $O/r02d_timeline_c2_full.txt > /dev/null 2> $O/r02d_tl.err || tail -3 $O/r02d_tl.err
cut -c1-110 $O/r02d_timeline_c2_full.txt
for env in "X=1" "TIC_PDL_CHAINS=1"; do
  echo "$env"; env $env timeout 120 python scripts/timeline.py --replays 400 --plain-only --variant full,itc,fusion 2>&1 | grep "^workload"
done
