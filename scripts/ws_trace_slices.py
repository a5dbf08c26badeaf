"""Skew between the slices of one role (debug): AMIRA_WS_TRACE=3 stamps of every slice of the role AMIRA_WS_TRACE_ROLE
(0 layer-0, 1 layer-1 input, 2 layer-1 recurrent, 3 joint, 4 vocabulary) for the units of M-tile 0; prints, per event, the
spread over the slices (relative to the earliest slice) and each slice's median lag."""
import os
import sys

import numpy as np

os.environ["AMIRA_WS_TRACE"] = "3"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import amira_b200 as A  # noqa: E402
from bench import encoded_len, make_workload  # noqa: E402

B = 1024
ctx = A.Context(device_id=0, decode_engine=4)
ctx.load_weights(A.synthetic_weights(3456))
_, _, lens = make_workload(B, 4567)
elens = np.array([encoded_len(int(x // 160 + 1)) for x in lens], np.int64)
T = int(elens.max())
g = torch.Generator(device="cuda")
g.manual_seed(2345)
enc = torch.randn((B, 1024, T), generator=g, device="cuda", dtype=torch.float32) * 0.5
tok = torch.zeros((B, 200), dtype=torch.int32, device="cuda")
nt = torch.zeros(B, dtype=torch.int32, device="cuda")
ns = torch.zeros(B, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.greedy_decode_raw(enc.data_ptr(), B, T, elens, tok.data_ptr(), nt.data_ptr(), ns.data_ptr())
torch.cuda.synchronize()
raw = ctx.debug_ws_trace(512)
print("SM of each CTA (by block index):", " ".join(str(int(v)) for v in raw.reshape(512, 256)[510, :148]))
tr = raw.astype(np.float64).reshape(512, 4, 64)  # [its][event][slice]
role = int(os.environ.get("AMIRA_WS_TRACE_ROLE", "1"))
nsl = [40, 40, 40, 10, 18][role]
names = ["dependency seen", "popped by the epilogue", "stores issued", "published"]
its = range(50, 200)
for e in range(4):
    x = np.array([tr[i, e, :nsl] for i in its if (tr[i, e, :nsl] > 0).all()])
    if not len(x):
        print(f"{names[e]}: no data")
        continue
    lag = x - x.min(axis=1, keepdims=True)
    print(f"{names[e]:>24}: spread over slices median {np.median(lag.max(axis=1)) / 1e3:.2f} us, max {lag.max() / 1e3:.2f} us; "
          f"median lag by slice (us): " + " ".join(f"{v / 1e3:.1f}" for v in np.median(lag, axis=0)))
x1 = np.array([tr[i, 1, :nsl] for i in its])
x3 = np.array([tr[i, 3, :nsl] for i in its])
ok = (x1 > 0).all(axis=1) & (x3 > 0).all(axis=1)
print(f"popped (earliest slice) -> published (last slice): median {np.median((x3.max(axis=1) - x1.min(axis=1))[ok]) / 1e3:.2f} us; "
      f"per slice popped -> published: median {np.median((x3 - x1)[ok]) / 1e3:.2f} us")
per = np.diff(x3.max(axis=1)[ok])
print(f"tick period (last slice published): median {np.median(per) / 1e3:.2f} us")
