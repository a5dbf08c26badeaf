#!/bin/bash
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-stream"
$CMD > gpurun_out/plain_fe.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fe_logmel" -s 1 -c 1 -o gpurun_out/prof_fe $CMD > gpurun_out/ncu_fe.log 2>&1
tail -2 gpurun_out/ncu_fe.log
