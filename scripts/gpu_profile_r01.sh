#!/bin/bash
# Round-1 profiling pass on ONE B200 (run under gpurun): launch lists of the bench commands, the c2 DRAM-traffic pass, and one
# `ncu --set full` capture per hot kernel.  Every ncu run is preceded by the same command WITHOUT ncu (must exit 0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
run_plain() { python bench.py --profile --steps 2 --warmup 3 "$@" > $O/plain.log 2>&1 || { echo "plain run failed: $*"; tail -5 $O/plain.log; exit 1; }; }

# ---- launch lists (gpu__time_duration.sum) of the bench command itself (CUDA-graph replay included)
for w in c2 c4 itc:16384x768; do
  n=${w//:/_}
  run_plain --workload $w
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_$n.csv python bench.py --profile --steps 2 --warmup 3 --workload $w > $O/ncu_l_$n.log 2>&1
  echo "launch list $w: $(grep -c umma_gemm_kernel $O/launches_$n.csv) tcgen05 launches"
done
# ---- c2: DRAM traffic of every launch of a step (eager launches so that the order is the program order)
run_plain --workload c2 --no-graph
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 400 --csv --log-file $O/launches_c2_traffic.csv \
  python bench.py --profile --no-graph --steps 2 --warmup 3 --workload c2 > $O/ncu_t_c2.log 2>&1
# ---- full captures
P16="--profile --no-graph --steps 2 --warmup 3 --workload itc:16384x768"
run_plain --workload itc:16384x768 --no-graph
$NCU --set full --import-source on -k regex:ItcFwdEpi -s 3 -c 1 -o $O/prof_itc_fwd_16384x768 python bench.py $P16 > $O/ncu_f1.log 2>&1
$NCU --set full --import-source on -k regex:ItcBwdEpi -s 3 -c 1 -o $O/prof_itc_bwd_16384x768 python bench.py $P16 > $O/ncu_f2.log 2>&1
$NCU --set full --import-source on -k regex:StoreEpi -s 6 -c 2 -o $O/prof_gemm_dtdv_16384x768 python bench.py $P16 > $O/ncu_f3.log 2>&1
PC4="--profile --no-graph --steps 2 --warmup 3 --workload c4"
run_plain --workload c4 --no-graph
$NCU --set full --import-source on -k "regex:attn_pool_mma_kernel<2, 0" -s 3 -c 1 -o $O/prof_attn_fwd_4096 python bench.py $PC4 > $O/ncu_f4.log 2>&1
$NCU --set full --import-source on -k "regex:attn_pool_mma_kernel<2, 1" -s 3 -c 1 -o $O/prof_attn_bwd_4096 python bench.py $PC4 > $O/ncu_f5.log 2>&1
tail -1 $O/ncu_f1.log $O/ncu_f2.log $O/ncu_f3.log $O/ncu_f4.log $O/ncu_f5.log
ls -la $O/*.ncu-rep
