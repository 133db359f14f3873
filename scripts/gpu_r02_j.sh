#!/bin/bash
# 8 GPUs: timeline of the peer-memory c2 step + bench lines c2 / c3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 scripts/prof_timeline_dist.py --workload c2 --out $O/r02j_timeline_c2_g8.txt > /dev/null 2> $O/r02j_tl_c2.err || tail -5 $O/r02j_tl_c2.err
cut -c1-130 $O/r02j_timeline_c2_g8.txt
timeout 300 $TR --master-port 29552 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > $O/r02j_bench_c2_g8.json 2> $O/r02j_bench_c2_g8.err; echo "bench c2 g8 rc=$?"; tail -2 $O/r02j_bench_c2_g8.err
timeout 300 $TR --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 5 --workload c3 --no-cpu-baseline > $O/r02j_bench_c3_g8.json 2> $O/r02j_bench_c3_g8.err; echo "bench c3 g8 rc=$?"; tail -2 $O/r02j_bench_c3_g8.err
python - <<'PY'
import json
for n in ("c2_g8","c3_g8"):
    try:
        d=json.loads(open("gpurun_out/r02j_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("global_loss_check"))
        for k in d.get("kernels",[]): print("   %-90s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:90],k["ms"],k["achieved"],k["unit"],k["frac"]))
    except Exception as e:
        print(n, "parse failed", e)
PY
