"""Times bench.py's stream-group leg alone (cfg4 through the orchestrator, incremental mode)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amira_b200 as A
import bench
ctx = A.Context(device_id=0)
ctx.load_weights(A.synthetic_weights(3456))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
print(json.dumps(bench.run_stream_group(A, ctx, n, 40, 5)))
