"""BASELINE cfg0: a single 10 s utterance through the library (host buffers), latency per call."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import amira_b200 as A
from conftest import synth_pcm
ctx = A.Context(device_id=0); ctx.load_weights(A.synthetic_weights(3456))
pcm = synth_pcm(10.0, 1)
offs = np.array([0, pcm.size], np.int64)
rng = np.random.default_rng(0)
enc = (0.5 * rng.standard_normal((1, 1024, 126))).astype(np.float32)
def t(fn, k=50):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    for _ in range(k): fn()
    return (time.perf_counter() - t0) / k * 1e3
print("front end 10 s, host buffers: %.3f ms" % t(lambda: ctx.preprocess_pcm16(pcm, offs)))
st = A.DecoderState.new(1)
print("greedy decode T=126, host buffers, with state: %.3f ms" % t(lambda: ctx.greedy_decode(enc, [126], state=st)))
toks, _, steps = ctx.greedy_decode(enc, [126])
print("steps", int(steps[0]), "tokens", len(toks[0]))
ctx.profile(True)
ctx.greedy_decode(enc, [126]); ctx.preprocess_pcm16(pcm, offs)
print({k: ctx.kernel_ms(k) for k in ("fe_logmel", "fe_normalize", "enc_proj", "greedy")})
