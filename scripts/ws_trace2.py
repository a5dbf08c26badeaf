"""Inside one role's epilogue (debug): AMIRA_WS_TRACE=2 stamps of slice 0 of the role AMIRA_WS_TRACE_ROLE (default 1 = layer-1
input half) for every unit; prints the median duration of each section and the unit-to-unit period."""
import os
import sys

import numpy as np

os.environ["AMIRA_WS_TRACE"] = "2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import amira_b200 as A  # noqa: E402
from bench import encoded_len, make_workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
tr = ctx.debug_ws_trace(512).astype(np.float64)  # [its][8][32]
nt_ = int((tr[8, :, 0] > 0).sum())
role = int(os.environ.get("AMIRA_WS_TRACE_ROLE", "1"))
names0 = ["unit popped", "accumulator full", "drained", "decisions here", "loads here + gathered", "tile released", "-",
          "arithmetic done, stores issued", "signal thread: all arrived", "signal thread: published"]
names = names0 if role == 0 else ["unit popped", "control row here", "partner flag seen", "partial sums + cell state here", "accumulator full", "drained",
         "gathered", "arithmetic done, stores issued", "signal thread: all arrived", "signal thread: published"]
lo, hi = 50, 400
print("M-tiles:", nt_)
for mt in range(nt_):
    rows = [i for i in range(lo, hi) if tr[i, mt, 0] > 0 and tr[i, mt, 9] > 0]
    print(f"tile {mt}: sections (us, median) from 'unit popped':")
    for k in range(1, 10):
        v = [tr[i, mt, k] - tr[i, mt, 0] for i in rows if tr[i, mt, k] > 0]
        print(f"   {names[k]:>34}: {np.median(v) / 1e3:6.2f}")
# unit-to-unit period in program order: (it, mt) sorted by pop time
ev = sorted((tr[i, mt, 0], i, mt) for i in range(lo, hi) for mt in range(nt_) if tr[i, mt, 0] > 0)
d = np.diff([e[0] for e in ev])
print(f"unit period (pop to pop): median {np.median(d) / 1e3:.2f} us, mean {np.mean(d) / 1e3:.2f} us")
fin = sorted((tr[i, mt, 7], tr[i, mt, 0]) for i in range(lo, hi) for mt in range(nt_) if tr[i, mt, 0] > 0 and tr[i, mt, 7] > 0)
gap = [fin[k + 1][1] - fin[k][0] for k in range(len(fin) - 1)]
print(f"stores issued -> next unit popped: median {np.median(gap) / 1e3:.2f} us")
