#!/bin/bash
# usage: gpu_r02_p.sh N [full]   — c5 sweep at N GPUs; with "full" also the c2 / c3 bench lines and the c2 timeline
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29571 scripts/sweep_c5.py > gpurun_out/c5_g$N.log 2>&1; echo sweep rc=$?; grep "B=" gpurun_out/c5_g$N.log | tail -20
if [ "$2" = "full" ]; then
  timeout 600 $TR --master-port 29572 bench.py --gpus $N > gpurun_out/r02p_c2_g$N.json 2> gpurun_out/r02p_c2_g$N.err; echo c2 rc=$?
  timeout 600 $TR --master-port 29573 bench.py --gpus $N --workload c3 > gpurun_out/r02p_c3_g$N.json 2> gpurun_out/r02p_c3_g$N.err; echo c3 rc=$?
  timeout 600 $TR --master-port 29574 scripts/prof_timeline_dist.py --workload c2 --out gpurun_out/r02p_timeline_c2_g$N.txt > gpurun_out/r02p_tl.err 2>&1; echo tl rc=$?
  timeout 600 $TR --master-port 29575 scripts/prof_timeline_dist.py --workload c3 --out gpurun_out/r02p_timeline_c3_g$N.txt > gpurun_out/r02p_tl3.err 2>&1; echo tl3 rc=$?
  python - <<PY
import json
for w in ("c2","c3"):
    try:
        d=json.loads(open("gpurun_out/r02p_%s_g$N.json"%w).read().strip().splitlines()[-1])
        print(w,"N=$N ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"],d["value"],d["e2e"]["value"]), d.get("global_loss_check"))
    except Exception as e: print(w,"failed",e)
PY
  head -45 gpurun_out/r02p_timeline_c2_g$N.txt | cut -c1-130
fi
