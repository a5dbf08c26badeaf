#!/bin/bash
# Round-1 profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture of the decode kernel.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-stream"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"greedy_ws|fe_logmel" -s 2 -c 2 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
