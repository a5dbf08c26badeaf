#!/bin/bash
for i in 1 2; do for lib in libamira_b200_base.so libamira_b200.so; do
  AMIRA_B200_LIB=$lib python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-stream 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['kernel_ms_per_step'])"
done; done
