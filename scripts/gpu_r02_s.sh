#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/r02s_pytest_dist_$i.log 2>&1; echo "pytest dist $i rc=$?"; tail -3 gpurun_out/r02s_pytest_dist_$i.log
done
