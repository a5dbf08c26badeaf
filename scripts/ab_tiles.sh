#!/bin/bash
# A/B of the lane plan of the decode kernel: "tiles spec" pairs — AMIRA_WS_TILES = number of M-tiles the batch's streams share,
# AMIRA_WS_SPEC = speculation depth by live M-tiles ("auto" = the planner's choice).  Prints the decode kernel's ms per step on
# the bench workload and the cfg3 stand-alone times.
cd "$(dirname "$0")/.."
IFS=';' read -ra CFGS <<< "${AB_CFGS:-auto auto;3 1,1,1;4 1,1,1,1;5 1,1,1,1,1;4 0;6 0;8 0}"
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  if [ "$1" = auto ]; then unset AMIRA_WS_TILES; else export AMIRA_WS_TILES=$1; fi
  if [ "$2" = auto ]; then unset AMIRA_WS_SPEC; else export AMIRA_WS_SPEC=$2; fi
  timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
x=d['extra']
print('tiles=$1 spec=$2 greedy %.3f ms  step %.3f ms  cfg3 T126 %.3f ms  T376 %.3f ms' % (d['kernel_ms_per_step']['greedy'], d['ms_per_step'], x['cfg3_greedy_256xT126']['ms'], x['cfg3_greedy_256xT376']['ms']))"
done
