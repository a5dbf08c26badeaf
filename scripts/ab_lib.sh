#!/bin/bash
# A/B on one GPU between shared libraries: ab_lib.sh libA.so libB.so ...  (decode kernel ms of bench.py's workload, two rounds)
for i in 1 2; do for lib in "$@"; do
  AMIRA_B200_LIB=$lib python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['kernel_ms_per_step'], d['config']['tokens_per_step'])"
done; done
