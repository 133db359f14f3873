#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_c5.py > gpurun_out/c5_g1.log 2>&1; echo sweep rc=$?; grep -E "B= +(4096|8192)" gpurun_out/c5_g1.log
timeout 300 python scripts/ncu_step.py --workload itc:16384x256 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:umma_gemm -c 4 -f -o gpurun_out/r02_ncu_itc16k_d256 \
   python scripts/ncu_step.py --workload itc:16384x256 > gpurun_out/r02o_ncu.log 2>&1; echo ncu rc=$?; tail -3 gpurun_out/r02o_ncu.log
ls -la gpurun_out/*.ncu-rep
