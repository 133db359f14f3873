#!/bin/bash
# 2 GPUs: multi-GPU parity tests + timelines of the peer-memory c2 / c3 steps + bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/r02i_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 $O/r02i_pytest_dist.log | cut -c1-200
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 scripts/prof_timeline_dist.py --workload c2 --out $O/r02i_timeline_c2_g2.txt > /dev/null 2> $O/r02i_tl_c2.err || tail -5 $O/r02i_tl_c2.err
cut -c1-130 $O/r02i_timeline_c2_g2.txt
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 50 --warmup 5 > $O/r02i_bench_c2_g2.json 2> $O/r02i_bench_c2_g2.err; echo "bench c2 g2 rc=$?"; tail -2 $O/r02i_bench_c2_g2.err
timeout 300 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --workload c3 > $O/r02i_bench_c3_g2.json 2> $O/r02i_bench_c3_g2.err; echo "bench c3 g2 rc=$?"; tail -2 $O/r02i_bench_c3_g2.err
python - <<'PY'
import json
for n in ("c2_g2","c3_g2"):
    try:
        d=json.loads(open("gpurun_out/r02i_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("global_loss_check"))
        for k in d.get("kernels",[]): print("   %-90s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:90],k["ms"],k["achieved"],k["unit"],k["frac"]))
    except Exception as e:
        print(n, "parse failed", e)
PY
