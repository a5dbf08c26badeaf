"""Per-step critical-path breakdown of the weight-stationary decode engine (debug): runs the bench workload's decode once
with AMIRA_WS_TRACE=1 and prints, for M-tile 0, the median time between consecutive events of one decode step."""
import os
import sys

import numpy as np

os.environ["AMIRA_WS_TRACE"] = "1"
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
tr = ctx.debug_ws_trace(512).astype(np.float64)
names = {0: "A", 1: "BI", 2: "BH", 3: "C", 4: "D"}
ev = {0: "dep seen", 1: "1st TMA landed", 2: "MMAs issued", 3: "acc full seen", 4: "signalled", 5: "ctl seen (A)"}
its = [i for i in range(4, 500) if tr[i, 30] > 0 and tr[i - 1, 30] > 0]
print("steps traced", len(its), "median step period (ctl->ctl) us:", np.median([(tr[i, 30] - tr[i - 1, 30]) for i in its]) / 1e3)
for lo, hi in ((4, 120), (140, 300), (380, 470)):
    sel = [i for i in its if lo <= i < hi]
    if not sel:
        continue
    print(f"--- steps {lo}..{hi}: period {np.median([(tr[i, 30] - tr[i - 1, 30]) for i in sel]) / 1e3:.2f} us; offsets from previous ctl (us):")
    for role in range(5):
        row = []
        for e in range(6):
            v = [tr[i, role * 6 + e] - tr[i - 1, 30] for i in sel if tr[i, role * 6 + e] > 0]
            row.append(f"{ev[e]}={np.median(v) / 1e3:7.2f}" if v else f"{ev[e]}=   --  ")
        print(f"  {names[role]:>2}: " + "  ".join(row))
    v = [tr[i, 30] - tr[i - 1, 30] for i in sel]
    print(f"  ctl done = {np.median(v) / 1e3:7.2f};  D-role MMA thread waited on TMA data {np.median([tr[i, 31] for i in sel]) / 1.965e3:.2f} us per unit (after the first chunk)")
