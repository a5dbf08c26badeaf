"""Decode latency of both engines at small batch sizes (device-resident inputs); AMIRA_WS_SPEC selects the speculation depth."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
blob = A.synthetic_weights(3456)
for eng in (1, 4):
    ctx = A.Context(device_id=0, decode_engine=eng); ctx.load_weights(blob)
    for B in (1, 4, 16, 64):
        T = 126
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        enc = torch.randn((B, 1024, T), generator=g, device="cuda") * 0.5
        tok = torch.zeros((B, 200), dtype=torch.int32, device="cuda"); nt = torch.zeros(B, dtype=torch.int32, device="cuda"); ns = torch.zeros(B, dtype=torch.int32, device="cuda")
        for _ in range(2):
            ctx.greedy_decode_raw(enc.data_ptr(), B, T, None, tok.data_ptr(), nt.data_ptr(), ns.data_ptr())
        ctx.profile(True)
        for _ in range(3):
            ctx.greedy_decode_raw(enc.data_ptr(), B, T, None, tok.data_ptr(), nt.data_ptr(), ns.data_ptr())
        ms, n = ctx.kernel_ms("greedy"); ctx.profile(False)
        print(f"engine {eng} B={B}: greedy {ms / n:.3f} ms, steps max {int(ns.max())} -> {1e3 * ms / n / max(int(ns.max()), 1):.1f} us/step")
    ctx.close()
