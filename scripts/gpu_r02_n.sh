#!/bin/bash
# c5 points: kernel timelines at the sizes where the sweep is off its trend
mkdir -p gpurun_out
for w in itc:4096x256 itc:8192x256 itc:16384x256 itc:8192x768 itc:16384x768 itc:65536x768; do
  timeout 300 python scripts/prof_timeline.py --workload $w --replays 2 --out gpurun_out/r02n_tl_${w/:/_}.txt > gpurun_out/r02n_tl.err 2>&1 || tail -5 gpurun_out/r02n_tl.err
  head -40 gpurun_out/r02n_tl_${w/:/_}.txt | cut -c1-150
done
