import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
n, chunk, T = 1024, 2560, 3
ctx = A.Context(device_id=0); ctx.load_weights(A.synthetic_weights(3456))
rng = np.random.default_rng(99)
pcm = torch.from_numpy((rng.standard_normal(n * chunk) * 3000).astype(np.int16)).pin_memory()
offsets = np.arange(n + 1, dtype=np.int64) * chunk
feats = torch.empty((n, 128, 32), dtype=torch.float32).pin_memory()
enc = torch.from_numpy((0.5 * rng.standard_normal((n, 1024, T))).astype(np.float32)).pin_memory()
tok = torch.zeros((n, 200), dtype=torch.int32).pin_memory(); ntok = torch.zeros(n, dtype=torch.int32).pin_memory()
nst = torch.zeros(n, dtype=torch.int32).pin_memory()
flens = np.zeros(n, np.int64)
slots = np.array([ctx.stream_open() for _ in range(n)], np.int32)
enc_d = enc.cuda(); pcm_d = pcm.cuda(); feats_d = feats.cuda(); tok_d = tok.cuda(); ntok_d = ntok.cuda(); nst_d = nst.cuda()
def t(fn, k=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
print("fe host    ", t(lambda: ctx.preprocess_pcm16_raw(pcm.data_ptr(), offsets, n, feats.data_ptr(), 32, flens)))
print("fe device  ", t(lambda: ctx.preprocess_pcm16_raw(pcm_d.data_ptr(), offsets, n, feats_d.data_ptr(), 32, flens)))
print("dec host   ", t(lambda: ctx.stream_decode_raw(slots, enc.data_ptr(), T, None, tok.data_ptr(), ntok.data_ptr(), nst.data_ptr())))
print("dec device ", t(lambda: ctx.stream_decode_raw(slots, enc_d.data_ptr(), T, None, tok_d.data_ptr(), ntok_d.data_ptr(), nst_d.data_ptr())))
print("steps max/mean", int(nst_d.max()), float(nst_d.float().mean()))
ctx.profile(True)
ctx.stream_decode_raw(slots, enc_d.data_ptr(), T, None, tok_d.data_ptr(), ntok_d.data_ptr(), nst_d.data_ptr())
ctx.preprocess_pcm16_raw(pcm_d.data_ptr(), offsets, n, feats_d.data_ptr(), 32, flens)
print({k: ctx.kernel_ms(k) for k in ("fe_logmel", "enc_proj", "greedy")})
