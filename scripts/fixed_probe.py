"""Fixed cost of one greedy-decode launch: kernel ms against the longest stream's step count, for several T and B."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
ctx = A.Context(device_id=0); ctx.load_weights(A.synthetic_weights(3456))
for B in (128, 1024):
    for T in (1, 2, 4, 8, 16, 32, 64):
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        enc = torch.randn((B, 1024, T), generator=g, device="cuda") * 0.5
        tok = torch.zeros((B, 200), dtype=torch.int32, device="cuda"); nt = torch.zeros(B, dtype=torch.int32, device="cuda"); ns = torch.zeros(B, dtype=torch.int32, device="cuda")
        for _ in range(3):
            ctx.greedy_decode_raw(enc.data_ptr(), B, T, None, tok.data_ptr(), nt.data_ptr(), ns.data_ptr())
        ctx.profile(True)
        for _ in range(5):
            ctx.greedy_decode_raw(enc.data_ptr(), B, T, None, tok.data_ptr(), nt.data_ptr(), ns.data_ptr())
        ms, n = ctx.kernel_ms("greedy"); pm, pn = ctx.kernel_ms("enc_proj")
        ctx.profile(False)
        print(f"B={B} T={T}: greedy {ms / n:.3f} ms, enc_proj {pm / pn:.3f} ms, steps max {int(ns.max())} mean {float(ns.float().mean()):.1f}")
