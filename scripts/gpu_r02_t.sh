#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for i in 1 2; do
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/r02t_pytest_dist_$i.log 2>&1; echo "pytest dist $i rc=$?"; tail -2 gpurun_out/r02t_pytest_dist_$i.log
done
TIC_PDL_CHAINS=1 timeout 300 $TR --master-port 29582 tests/dist_gpu_check.py --mode peer --graph --live > gpurun_out/r02t_pdl_graph.log 2>&1; echo "pdl graph live rc=$?"; grep "dist check\|tic:" gpurun_out/r02t_pdl_graph.log | cut -c1-250
timeout 600 $TR --master-port 29572 bench.py --gpus $N > gpurun_out/r02t_c2_g$N.json 2> gpurun_out/r02t_c2_g$N.err; echo c2 rc=$?
TIC_PDL_CHAINS=1 timeout 600 $TR --master-port 29576 bench.py --gpus $N > gpurun_out/r02t_c2_g${N}_pdl.json 2> gpurun_out/r02t_c2_g${N}_pdl.err; echo c2pdl rc=$?
TIC_PDL_CHAINS=1 timeout 600 $TR --master-port 29574 scripts/prof_timeline_dist.py --workload c2 --out gpurun_out/r02t_timeline_c2_g${N}_pdl.txt > gpurun_out/r02t_tl.err 2>&1; echo tl rc=$?
python - <<PY
import json
for w in ("c2_g$N","c2_g${N}_pdl"):
    try:
        d=json.loads(open("gpurun_out/r02t_%s.json"%w).read().strip().splitlines()[-1])
        print(w,"ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"],d["value"],d["e2e"]["value"]), d.get("global_loss_check"))
    except Exception as e: print(w,"failed",e)
PY
head -44 gpurun_out/r02t_timeline_c2_g${N}_pdl.txt | cut -c1-130
