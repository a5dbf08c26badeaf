#!/bin/bash
# Profiling recipe, part 1 (B200_PROFILING.md): plain run first, then the per-launch duration list of the same command.
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-stream"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-400
