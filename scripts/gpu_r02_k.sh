#!/bin/bash
# 2 GPUs: push-form exchange: parity tests, timeline, bench c2 (push vs pull), c3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/r02k_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -5 $O/r02k_pytest_dist.log | cut -c1-250
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 scripts/prof_timeline_dist.py --workload c2 --out $O/r02k_timeline_c2_g2.txt > /dev/null 2> $O/r02k_tl_c2.err || tail -5 $O/r02k_tl_c2.err
cut -c1-130 $O/r02k_timeline_c2_g2.txt
for env in "X=1" "TIC_PEER_PUSH=0" "TIC_CODE_WARM=0"; do
  env $env timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > $O/r02k_b.json 2> $O/r02k_b.err || tail -3 $O/r02k_b.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02k_b.json').read().strip().splitlines()[-1]); print('$env', 'c2 g2 ms/step %.4f value %.3e e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']), d.get('global_loss_check',{}).get('rel_err'))"
done
