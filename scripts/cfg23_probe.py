"""Stand-alone numbers for BASELINE configs 2 and 3 (documentation only; bench.py measures config 5):
cfg2 = front end only, 64 x 30 s PCM resident on the device; cfg3 = greedy decode only, 256 streams x T = 126."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
from bench import make_workload

ctx = A.Context(device_id=0); ctx.load_weights(A.synthetic_weights(3456))
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(n): fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
# cfg2
pcm, offsets, lens = make_workload(64, 1234, 30.0, 30.0)
pcm_d = torch.from_numpy(pcm).cuda(); t_stride = 3008
feats = torch.empty((64, 128, t_stride), dtype=torch.float32, device="cuda"); fl = np.zeros(64, np.int64)
ms = timed(lambda: ctx.preprocess_pcm16_raw(pcm_d.data_ptr(), offsets, 64, feats.data_ptr(), t_stride, fl))
nbytes = 2 * lens.sum() + 4 * 128 * fl.sum()
print(f"cfg2 front end 64 x 30 s: {ms:.3f} ms, {lens.sum() / 16000 / ms * 1e3:.0f} audio-s/s, {nbytes / ms / 1e6:.0f} GB/s algorithmic ({nbytes / ms / 1e6 / 6549.4:.3f} of measured HBM)")
# cfg3
g = torch.Generator(device="cuda"); g.manual_seed(2345)
for B, T in ((256, 126), (256, 376)):
    enc = torch.randn((B, 1024, T), generator=g, device="cuda") * 0.5
    tok = torch.zeros((B, 200), dtype=torch.int32, device="cuda"); nt = torch.zeros(B, dtype=torch.int32, device="cuda"); ns = torch.zeros(B, dtype=torch.int32, device="cuda")
    lens64 = np.full(B, T, np.int64)
    ms = timed(lambda: ctx.greedy_decode_raw(enc.data_ptr(), B, T, lens64, tok.data_ptr(), nt.data_ptr(), ns.data_ptr()), 5)
    steps = int(ns.sum()); fl_ = steps * 15_244_800 + B * T * 1_310_720
    print(f"cfg3 decode {B} x T={T}: {ms:.2f} ms, {B * T * 0.08 / ms * 1e3:.0f} audio-s/s, {steps} steps ({int(ns.max())} max), {fl_ / ms / 1e9:.1f} TFLOP/s algorithmic")
