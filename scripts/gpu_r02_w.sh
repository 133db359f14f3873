#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_new_entry_points.py -q -m gpu > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02w_pytest.log
timeout 600 $TR --master-port 29572 bench.py --gpus 2 > gpurun_out/r02w_c2_g2.json 2> gpurun_out/r02w_c2_g2.err; echo c2 rc=$?
timeout 600 $TR --master-port 29574 scripts/prof_timeline_dist.py --workload c2 --out gpurun_out/r02w_timeline_c2_g2.txt > gpurun_out/r02w_tl.err 2>&1; echo tl rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r02w_c2_g2.json").read().strip().splitlines()[-1])
print("c2 g2 ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"],d["value"],d["e2e"]["value"]), d.get("global_loss_check"))
PY
grep -n "lse_rows\|peer_push\|peer_signal\|Fill" gpurun_out/r02w_timeline_c2_g2.txt | cut -c1-110
