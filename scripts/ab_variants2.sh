#!/bin/bash
# A/B on one GPU: base library, then the current library under each AMIRA_WS_VARIANT given
run() { python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['kernel_ms_per_step']['greedy'], d['config']['tokens_per_step'])"; }
AMIRA_B200_LIB=libamira_b200_base.so run base
for v in "$@"; do AMIRA_WS_VARIANT=$v run "variant $v"; done
AMIRA_B200_LIB=libamira_b200_base.so run base
