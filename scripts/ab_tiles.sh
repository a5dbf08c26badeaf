#!/bin/bash
# A/B of the lane plan of the decode kernel: AMIRA_WS_TILES = number of M-tiles the batch's streams are packed into
# (unset = the planner's choice).  Prints the decode kernel's ms per step on the bench workload and the cfg3 stand-alone times.
cd "$(dirname "$0")/.."
for cfg in "" ${AB_TILES:-3 4 5 6 7 8}; do
  AMIRA_WS_TILES=$cfg timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
x=d['extra']
print('tiles=${cfg:-auto} spec=${AMIRA_WS_SPEC:-default} greedy %.3f ms  step %.3f ms  cfg3 T126 %.3f ms  T376 %.3f ms' % (d['kernel_ms_per_step']['greedy'], d['ms_per_step'], x['cfg3_greedy_256xT126']['ms'], x['cfg3_greedy_256xT376']['ms']))"
done
