#!/bin/bash
# Profiling recipe, part 2: plain run first, then ONE full capture of one launch each of the decode kernel and the front end.
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-stream"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"greedy_ws|fe_logmel" -s 2 -c 2 -o gpurun_out/prof_r1e $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
