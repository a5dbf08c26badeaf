#!/bin/bash
# A/B timing of AMIRA_WS_VARIANT code paths on one GPU: decode kernel ms of bench.py's workload.
for v in "$@"; do
  AMIRA_WS_VARIANT=$v python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', d['kernel_ms_per_step']['greedy'], d['config']['tokens_per_step'])"
done
