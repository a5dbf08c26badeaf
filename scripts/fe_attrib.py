"""Timing attribution of the fused front-end kernel: AMIRA_FE_DEBUG bit mask (1 no normalisation, 2 no mel phase, 8 no transforms)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
ctx = A.Context(device_id=0)
B, secs = 1024, 17
n = secs * 16000
pcm = torch.randint(-3000, 3000, (B * n,), dtype=torch.int16, device="cuda")
offs = np.arange(B + 1, dtype=np.int64) * n
L = n // 160 + 1
foff = np.arange(B + 1, dtype=np.int64) * (128 * L)
out = torch.empty((B * 128 * L,), dtype=torch.float32, device="cuda")
lens = np.zeros(B, np.int64)
for flags in (0, 1, 2, 3, 8, 9, 10, 11):
    os.environ["AMIRA_FE_DEBUG"] = str(flags)
    for _ in range(2): ctx.preprocess_pcm16_packed_raw(pcm.data_ptr(), offs, B, out.data_ptr(), foff, lens)
    ctx.profile(True)
    for _ in range(3): ctx.preprocess_pcm16_packed_raw(pcm.data_ptr(), offs, B, out.data_ptr(), foff, lens)
    ms, k = ctx.kernel_ms("fe_logmel"); ctx.profile(False)
    print(f"flags={flags:2d}: fe kernel {ms / k:.3f} ms")
