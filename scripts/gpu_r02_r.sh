#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29581 tests/dist_gpu_check.py --mode peer --live > gpurun_out/r02r_live_eager.log 2>&1; echo "eager live rc=$?"; grep -v "Warning\|^W1018\|^\[rank" gpurun_out/r02r_live_eager.log | head -20 | cut -c1-400
timeout 300 $TR --master-port 29582 tests/dist_gpu_check.py --mode peer --graph --live > gpurun_out/r02r_live_graph.log 2>&1; echo "graph live rc=$?"; grep -v "Warning\|^W1018\|^\[rank" gpurun_out/r02r_live_graph.log | head -20 | cut -c1-400
