#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the end-of-round-2 kernels: plain run first, then the per-launch duration list, then ONE
# full capture of one launch each of the decode kernel (pair form) and the fused front end, all of the same command.
# (With the cooperative-launch attribute the pair form does not start under ncu — LaunchFailed at the profiled launch; the kernel
# carries its own grid barrier and is launched without the attribute.)
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-stream --no-extras"
$CMD > gpurun_out/r2c_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2c_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_bench_steps2.csv $CMD > gpurun_out/r2c_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"greedy_ws|fe_fused" -s 2 -c 2 -o gpurun_out/prof_r2c $CMD > gpurun_out/r2c_ncu_full.log 2>&1
ncu -i gpurun_out/prof_r2c.ncu-rep --page raw --csv > gpurun_out/r2c_ncu_full_raw.csv 2>/dev/null
grep -v "^==WARNING\|^$" gpurun_out/r2c_ncu_full.log | tail -4 | cut -c1-200; cut -c1-300 gpurun_out/r2c_plain.log | tail -1; wc -l gpurun_out/r2c_ncu_full_raw.csv
