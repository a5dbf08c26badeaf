#!/bin/bash
# stability soak: repeated decode benches and test runs; any non-zero exit or timeout is reported
fail=0
for i in $(seq 1 8); do
  timeout 120 python bench.py --steps 4 --warmup 2 --no-cpu --no-stream > gpurun_out/soak_$i.json 2> gpurun_out/soak_$i.err || { echo "bench run $i FAILED rc=$?"; fail=1; }
done
for i in 1 2; do
  timeout 300 python -m pytest tests -q -x -m gpu 2>&1 | tail -1
done
python - <<'PY'
import json, glob
v = []
for f in sorted(glob.glob("gpurun_out/soak_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); v.append((d["kernel_ms_per_step"]["greedy"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "unreadable", e)
print("greedy ms:", [x[0] for x in v]); print("e2e ms:", [round(x[1], 1) for x in v])
PY
echo "soak fail=$fail"
