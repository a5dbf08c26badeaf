"""A streaming tick's front-end call (1024 segments of ~2900 samples, host buffers): normalised vs un-normalised entry, pinned vs pageable."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
ctx = A.Context(device_id=0)
n = 1024
for seg in (2560, 2880):
    L = seg // 160 + 1
    offs = np.arange(n + 1, dtype=np.int64) * seg
    foff = np.arange(n + 1, dtype=np.int64) * (128 * L)
    flens = np.zeros(n, np.int64)
    pcm_np = (np.random.default_rng(0).standard_normal(n * seg) * 3000).astype(np.int16)
    for pinned in (True, False):
        pcm = torch.from_numpy(pcm_np.copy())
        feats = torch.empty(n * 128 * L, dtype=torch.float32)
        if pinned:
            pcm, feats = pcm.pin_memory(), feats.pin_memory()
        def t(fn, k=30):
            for _ in range(5): fn()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(k): fn()
            torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
        a = t(lambda: ctx.preprocess_pcm16_packed_raw(pcm.data_ptr(), offs, n, feats.data_ptr(), foff, flens))
        b = t(lambda: ctx.logmel_pcm16_packed_raw(pcm.data_ptr(), offs, n, feats.data_ptr(), foff, flens))
        print(f"seg {seg} pinned {pinned}: normalised {a:.3f} ms, un-normalised {b:.3f} ms")
