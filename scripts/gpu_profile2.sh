#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="--profile --no-graph --steps 2 --warmup 3 --workload itc:16384x768"
python bench.py $P > gpurun_out/plain_itc16k.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ItcFwdEpi -s 3 -c 1 -o gpurun_out/prof_itc_fwd python bench.py $P > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ItcBwdEpi -s 3 -c 1 -o gpurun_out/prof_itc_bwd python bench.py $P > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:StoreEpi -s 6 -c 1 -o gpurun_out/prof_gemm python bench.py $P > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full1.log gpurun_out/ncu_full2.log gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep
