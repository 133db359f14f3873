#!/bin/bash
# BASELINE config c5: aux-loss (ITC) sweep, batch 4k-64k x d 256-1024, on the GPUs of ONE box (run under gpurun [--gpus N]).
#   scripts/sweep_c5.sh [N_GPUS] [STEPS]        -> gpurun_out/c5_g<N>_<B>x<d>.json, one bench.py JSON line each
# B is the per-GPU row count (the global batch is N_GPUS x B); shapes whose [B, N*B] bf16 gradient operand would not fit in
# 150 GB are skipped.  The CPU leg is run once per d (at the smallest B) to keep the sweep short.
cd "$(dirname "$0")/.."
N=${1:-1}
STEPS=${2:-20}
mkdir -p gpurun_out
for d in 256 512 768 1024; do
  first=1
  for B in 4096 8192 16384 32768 65536; do
    bytes=$((B * B * N * 2))
    if [ $bytes -gt 150000000000 ]; then echo "skip B=$B d=$d N=$N (operand $((bytes / 1000000000)) GB)"; continue; fi
    extra="--no-scale-point"
    if [ $first -eq 0 ]; then extra="$extra --no-cpu-baseline"; fi
    first=0
    out=gpurun_out/c5_g${N}_${B}x${d}.json
    if [ "$N" -eq 1 ]; then
      timeout 600 python bench.py --workload itc:${B}x${d} --steps $STEPS --warmup 3 $extra > $out 2> ${out%.json}.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $N --workload itc:${B}x${d} --steps $STEPS --warmup 3 $extra > $out 2> ${out%.json}.err
    fi
    python - "$out" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-40s %8.3f ms  %10.3g samples/s  e2e %10.3g" % (d["config"]["workload"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
  done
done
