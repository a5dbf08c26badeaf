"""Debug: timeline of ONE CTA (slice 0 of role AMIRA_WS_TRACE_ROLE, default layer-1 input) across the M-tiles of a decode step."""
import os, sys
import numpy as np
os.environ["AMIRA_WS_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
from bench import encoded_len, make_workload
B = 1024
ctx = A.Context(device_id=0, decode_engine=4); ctx.load_weights(A.synthetic_weights(3456))
_, _, lens = make_workload(B, 4567)
elens = np.array([encoded_len(int(x // 160 + 1)) for x in lens], np.int64); T = int(elens.max())
g = torch.Generator(device="cuda"); g.manual_seed(2345)
enc = torch.randn((B, 1024, T), generator=g, device="cuda", dtype=torch.float32) * 0.5
tok = torch.zeros((B, 200), dtype=torch.int32, device="cuda"); nt = torch.zeros(B, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.greedy_decode_raw(enc.data_ptr(), B, T, elens, tok.data_ptr(), nt.data_ptr(), None)
torch.cuda.synchronize()
tr = ctx.debug_ws_trace(1536).astype(np.float64).reshape(-1)
t1 = tr[:512 * 32].reshape(512, 32); t2 = tr[512 * 32:].reshape(512, 8, 8)
ev = ["dep seen", "1st TMA", "MMAs issued", "acc seen", "signalled", "stores issued", "all arrived", "fence done"]
for it in (100, 420, 440, 441):
    base = t2[it, 0, 0]
    print(f"step {it}: (us relative to tile 0 dep seen); previous ctl of tile 0 at {(t1[it - 1, 30] - base) / 1e3:.2f}")
    for mt in range(8):
        print(f"   tile {mt}: " + "  ".join(f"{ev[e]}={(t2[it, mt, e] - base) / 1e3:7.2f}" for e in range(8)))
