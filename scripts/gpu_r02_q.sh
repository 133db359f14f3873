#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$2" != "notest" ]; then
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r02q_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/r02q_pytest_dist.log
fi
timeout 600 $TR --master-port 29574 scripts/prof_timeline_dist.py --workload c2 --out gpurun_out/r02q_timeline_c2_g$N.txt > gpurun_out/r02q_tl.err 2>&1; echo tl rc=$?
timeout 600 $TR --master-port 29572 bench.py --gpus $N > gpurun_out/r02q_c2_g$N.json 2> gpurun_out/r02q_c2_g$N.err; echo c2 rc=$?
TIC_BENCH_SNAPSHOT=1 timeout 600 $TR --master-port 29576 bench.py --gpus $N > gpurun_out/r02q_c2_g${N}_snapshot.json 2> gpurun_out/r02q_c2_g${N}_snap.err; echo c2snap rc=$?
python - <<PY
import json
for w in ("c2_g$N","c2_g${N}_snapshot"):
    try:
        d=json.loads(open("gpurun_out/r02q_%s.json"%w).read().strip().splitlines()[-1])
        print(w,"ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"],d["value"],d["e2e"]["value"]), d.get("global_loss_check"))
    except Exception as e: print(w,"failed",e)
PY
head -48 gpurun_out/r02q_timeline_c2_g$N.txt | cut -c1-130
