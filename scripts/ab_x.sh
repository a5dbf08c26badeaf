#!/bin/bash
# A/B of the decode kernel's experimental switches: AB_X="x1 x2 ..." values of AMIRA_WS_X (bit mask), for AMIRA_WS_PAIR in AB_PAIR
cd "$(dirname "$0")/.."
for pr in ${AB_PAIR:-0 1}; do for x in ${AB_X:-0}; do
  AMIRA_WS_PAIR=$pr AMIRA_WS_X=$x timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
x=d['extra']
print('pair=$pr x=$x greedy %.3f ms  step %.3f ms  cfg3 T126 %.3f ms  T376 %.3f ms  parity %s' % (d['kernel_ms_per_step']['greedy'], d['ms_per_step'], x['cfg3_greedy_256xT126']['ms'], x['cfg3_greedy_256xT376']['ms'], d.get('parity',{}).get('exact')))"
done; done
