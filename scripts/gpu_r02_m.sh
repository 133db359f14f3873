#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/r02m_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 $O/r02m_pytest_dist.log | cut -c1-200
timeout 200 $TR --master-port 29561 tests/dist_gpu_check.py --graph 2>&1 | grep "dist check" | cut -c1-400
timeout 300 $TR --master-port 29541 scripts/prof_timeline_dist.py --workload c2 --out $O/r02m_timeline_c2_g2.txt > /dev/null 2> $O/r02m_tl_c2.err || tail -5 $O/r02m_tl_c2.err
grep -E "un-profiled|peer_|ItcFwd|ItcBwd|lse_rows" $O/r02m_timeline_c2_g2.txt | cut -c1-110
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > $O/r02m_b.json 2> $O/r02m_b.err || tail -3 $O/r02m_b.err
python -c "
import json; d=json.loads(open('gpurun_out/r02m_b.json').read().strip().splitlines()[-1]); print('c2 g2 ms/step %.4f value %.3e e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']), d.get('global_loss_check',{}).get('rel_err'))"
timeout 600 python -m pytest tests/test_gpu_config1.py -m gpu -x -q -s 2>&1 | tail -5 | cut -c1-300
