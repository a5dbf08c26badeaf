#!/bin/bash
# one full capture (source-level stall samples included) of one front-end launch, after a plain run of the same command
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-stream"
$CMD > gpurun_out/plain_fe.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fe_logmel" -s 2 -c 1 -o gpurun_out/prof_fe2 -f $CMD > gpurun_out/ncu_fe.log 2>&1
tail -2 gpurun_out/ncu_fe.log
