#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi nvlink -gt d | head -30 > gpurun_out/r02_nvlink_probe.txt 2>&1
timeout 300 $TR --master-port 29621 scripts/nvlink_bytes.py --workload c2 --steps 2000 --out gpurun_out/r02_nvlink_c2_g2 > gpurun_out/r02z_c2.log 2>&1; echo "c2 rc=$?"; tail -2 gpurun_out/r02z_c2.log | cut -c1-200
timeout 300 $TR --master-port 29622 scripts/nvlink_bytes.py --workload c3 --steps 300 --out gpurun_out/r02_nvlink_c3_g2 > gpurun_out/r02z_c3.log 2>&1; echo "c3 rc=$?"; tail -2 gpurun_out/r02z_c3.log | cut -c1-200
head -12 gpurun_out/r02_nvlink_probe.txt
