"""Small end-to-end pass for compute-sanitizer (memcheck): ragged front end, packed decode with uneven lengths, padded decode."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import amira_b200 as A
from conftest import synth_pcm
rng = np.random.default_rng(1)
with A.Context(device_id=0) as ctx:
    ctx.load_weights(A.synthetic_weights(3456))
    pcms = [synth_pcm(float(s), 10 + i) for i, s in enumerate((0.31, 1.07, 0.0, 2.503, 0.005))]
    offs = np.zeros(len(pcms) + 1, np.int64); offs[1:] = np.cumsum([p.size for p in pcms])
    pcm = np.concatenate(pcms)
    feats, lens = ctx.preprocess_pcm16(pcm, offs)
    blocks, lens2 = ctx.preprocess_pcm16_packed(pcm, offs)
    assert all(np.array_equal(blocks[b], feats[b, :, :int(lens[b])]) for b in range(len(pcms)))
    enc = [(0.5 * rng.standard_normal((1024, t))).astype(np.float32) for t in (7, 1, 0, 13, 5, 9, 2)]
    toks, st, steps = ctx.greedy_decode_packed(enc)
    T = 13
    pad = np.zeros((len(enc), 1024, T), np.float32)
    for b, e in enumerate(enc):
        pad[b, :, :e.shape[1]] = e
    toks2, st2, steps2 = ctx.greedy_decode(pad, [e.shape[1] for e in enc])
    assert toks == toks2 and steps.tolist() == steps2.tolist()
    print("sanitize smoke ok", [len(t) for t in toks])
