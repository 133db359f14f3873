#!/bin/bash
# Runs the tcgen05 GEMM self-test matrix on a B200 (one process per case, each under its own timeout).
cd "$(dirname "$0")/.."
BIN="socialmedia-textimage-classification-auxlosses_b200/_build/selftest"
OUT=gpurun_out/selftest.log
mkdir -p gpurun_out
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $OUT 2>&1
run() { echo "--- $*" >> $OUT; timeout 60 $BIN "$@" >> $OUT 2>&1; echo "exit=$?" >> $OUT; }
# K-major x K-major first (the ITC / linear-forward form)
run 0 0 128 64 64
run 0 0 128 128 128
run 0 0 256 256 512
run 0 0 300 200 136
run 0 0 512 768 1536 0
run 0 0 512 768 1536 1
# MN-major B (dX = dY * W), MN-major A (weight gradients), both
run 0 1 128 64 64
run 0 1 256 256 512
run 0 1 300 200 136
run 1 0 128 64 64
run 1 0 256 256 512
run 1 0 300 200 136
run 1 1 128 64 64
run 1 1 256 256 512
run 1 1 300 200 136
run 1 1 768 1536 256
# big ones with timing
run 0 0 4096 4096 768 0 20
run 0 0 8192 8192 768 1 20
run 0 1 8192 768 8192 0 20
run 1 1 768 1536 8192 0 20
run 0 0 16384 16384 768 1 10
tail -100 $OUT
