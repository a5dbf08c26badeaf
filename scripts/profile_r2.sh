#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the per-launch duration list, then ONE full capture of one launch each
# of the decode kernel and the fused front end, all of the same command.
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-stream --no-extras"
$CMD > gpurun_out/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv $CMD > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"greedy_ws|fe_fused" -s 2 -c 2 -o gpurun_out/prof_r2 $CMD > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw.csv 2>/dev/null
tail -2 gpurun_out/r2_ncu_full.log; cut -c1-300 gpurun_out/r2_plain.log | tail -1
