import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
ctx = A.Context(device_id=0)
for B, secs in ((1, 30), (8, 30), (64, 30), (296, 30), (1024, 10), (1024, 30)):
    n = secs * 16000
    pcm = torch.randint(-3000, 3000, (B * n,), dtype=torch.int16, device="cuda")
    offs = np.arange(B + 1, dtype=np.int64) * n
    L = n // 160 + 1; ts = (L + 31) // 32 * 32
    out = torch.empty((B, 128, ts), dtype=torch.float32, device="cuda")
    lens = np.zeros(B, np.int64)
    for _ in range(2): ctx.preprocess_pcm16_raw(pcm.data_ptr(), offs, B, out.data_ptr(), ts, lens)
    ctx.profile(True)
    for _ in range(3): ctx.preprocess_pcm16_raw(pcm.data_ptr(), offs, B, out.data_ptr(), ts, lens)
    ms, k = ctx.kernel_ms("fe_logmel"); ctx.profile(False)
    print(f"B={B} x {secs}s: fe kernel {ms / k:.3f} ms")
    del pcm, out
