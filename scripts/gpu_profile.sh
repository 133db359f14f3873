#!/bin/bash
# ncu evidence for profiles/: launch lists (device time per launch) and full captures of the hot kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="--profile --no-graph --steps 2 --warmup 3"
python bench.py $P --workload c2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2.csv python bench.py $P --workload c2 > gpurun_out/ncu_c2.log 2>&1
python bench.py $P --workload c4 > gpurun_out/plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c4.csv python bench.py $P --workload c4 > gpurun_out/ncu_c4.log 2>&1
python bench.py $P --workload itc:16384x768 > gpurun_out/plain_itc16k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_itc16k.csv python bench.py $P --workload itc:16384x768 > gpurun_out/ncu_itc16k.log 2>&1
# full captures (one launch each, after warm-up launches of the same kernel)
ncu --set full --clock-control none --import-source on -k regex:ItcFwdEpi -s 3 -c 1 -o gpurun_out/prof_itc_fwd python bench.py $P --workload itc:16384x768 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ItcBwdEpi -s 3 -c 1 -o gpurun_out/prof_itc_bwd python bench.py $P --workload itc:16384x768 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:StoreEpi -s 6 -c 1 -o gpurun_out/prof_gemm python bench.py $P --workload itc:16384x768 > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_pool_fwd -s 3 -c 1 -o gpurun_out/prof_attn_fwd python bench.py $P --workload c4 > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out | tail -20
tail -3 gpurun_out/ncu_full1.log gpurun_out/ncu_c2.log
