#!/bin/bash
# A/B of the blank-speculation schedule of the decode kernel: AMIRA_WS_SPEC = depth while 1,2,3,... M-tiles are alive.
# Prints the decode kernel's ms per step on the bench workload and the cfg3 stand-alone times.
cd "$(dirname "$0")/.."
for cfg in "0" "3" "3,2" "3,2,1" "3,3" "3,3,1" "3,3,2" "3,2,2" "2,2,1" "3,2,1,1" "3,1,1"; do
  AMIRA_WS_SPEC=$cfg timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
x=d['extra']
print('spec=$cfg greedy %.3f ms  step %.3f ms  cfg3 T126 %.3f ms  T376 %.3f ms' % (d['kernel_ms_per_step']['greedy'], d['ms_per_step'], x['cfg3_greedy_256xT126']['ms'], x['cfg3_greedy_256xT376']['ms']))"
done
for cfg in "0" "1" "2" "3"; do echo "small batches, spec=$cfg"; AMIRA_WS_SPEC=$cfg python scripts/small_batch_probe.py 2>&1 | grep "engine 4"; done
