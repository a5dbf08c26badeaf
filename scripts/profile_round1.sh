set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --utterances 512"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'greedy_tc_kernel|fe_logmel_kernel' -s 6 -c 2 -o gpurun_out/prof_r1 -f $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/plain.log | cut -c1-400
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out/
