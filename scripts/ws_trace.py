"""Per-tick timeline of the weight-stationary decode engine (debug): runs the bench workload's decode once with
AMIRA_WS_TRACE=1 and prints, per M-tile, when slice 0 of every role handled the tile's unit of a tick (offsets from the
tick's start = the previous tick's control update of M-tile 0), plus the tick period.  AMIRA_WS_TILES / AMIRA_WS_SPEC apply."""
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
tr = ctx.debug_ws_trace(512).astype(np.float64)  # [its][8][32]
names = {0: "A", 1: "BI", 2: "BH", 3: "C", 4: "D"}
ev = {0: "dep", 1: "tma1", 2: "mma", 3: "acc", 4: "sig", 5: "epi"}
nt_ = int((tr[8, :, 30] > 0).sum())
print("M-tiles traced:", nt_)
print("ticks per M-tile:", [int(tr[511, m, 31]) for m in range(nt_)], " decode steps:", int(ns.sum().item()), " tokens:", int(nt.clamp(min=0).sum().item()),
      " longest stream (steps):", int(ns.max().item()))
print("lane-ticks by kind: step %d, redo step %d, copy (restore) %d, copy (waiting for results) %d, stream ended %d, load %d, idle %d"
      % tuple(int(tr[k, 0, 31]) for k in range(7)))
for a_, b_ in ((4, 50), (50, 100), (100, 200), (200, 300), (300, 400), (400, 511)):
    v = [tr[i, 0, 30] - tr[i - 1, 0, 30] for i in range(a_, b_) if tr[i, 0, 30] > 0 and tr[i - 1, 0, 30] > 0]
    if v:
        print(f"  ticks {a_}..{b_}: median period {np.median(v) / 1e3:.2f} us, mean {np.mean(v) / 1e3:.2f} us")
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20, 200)
its = [i for i in range(lo, hi) if tr[i, 0, 30] > 0 and tr[i - 1, 0, 30] > 0]
per = np.median([tr[i, 0, 30] - tr[i - 1, 0, 30] for i in its]) / 1e3
print(f"ticks {lo}..{hi}: tick period (M-tile 0, A control update to the next) = {per:.2f} us")
print("offsets (us, medians) from the previous tick's control update of M-tile 0; events: dep = scheduler saw the dependency,"
      " tma1 = first operand tile landed, mma = last MMA issued, acc = accumulator drained from, sig = unit published, epi = stores issued")
for mt in range(nt_):
    for role in range(5):
        row = []
        for e in range(6):
            v = [tr[i, mt, role * 6 + e] - tr[i - 1, 0, 30] for i in its if tr[i, mt, role * 6 + e] > 0]
            row.append(f"{ev[e]}={np.median(v) / 1e3:7.2f}" if v else f"{ev[e]}=   --  ")
        print(f"  tile {mt} {names[role]:>2}: " + "  ".join(row))
    v = [tr[i, mt, 30] - tr[i - 1, 0, 30] for i in its if tr[i, mt, 30] > 0]
    print(f"  tile {mt} control update done = {np.median(v) / 1e3:7.2f}")
