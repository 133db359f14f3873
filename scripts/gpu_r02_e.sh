#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/kernel_times.py > $O/r02e_kernel_times_c2.txt 2> $O/r02e_kt.err || tail -5 $O/r02e_kt.err
grep -A3 "fused ITC kernel phases" $O/r02e_kernel_times_c2.txt
