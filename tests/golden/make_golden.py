"""Generates the golden fixtures in this directory from the CPU oracle (the reference itself cannot run here: Rust +
absent ONNX models).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import amira_b200 as A  # noqa: E402  (host-side weight generator only; no GPU needed)
import oracle as O  # noqa: E402
from conftest import synth_pcm  # noqa: E402

pcm = synth_pcm(1.5, 4242)
feats, L = O.preprocess(pcm.astype(np.float32) / 32768.0, "f64")
np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), pcm=pcm, features=feats, features_len=L)

seed = 3456
model = O.Model(blob=A.synthetic_weights(seed))
rng = np.random.default_rng(7)
B, T = 4, 32
enc = (0.5 * rng.standard_normal((B, 1024, T))).astype(np.float16).astype(np.float32)  # fp16-representable: small file
lens = np.array([32, 19, 1, 27], np.int64)
r = O.greedy_decode_batch(model, enc, enc_lens=lens, threads=0)
np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), seed=seed, enc=enc.astype(np.float16), lens=lens,
                    tokens=r["tokens"], n_tokens=r["n_tokens"], n_steps=r["n_steps"], min_margin=r["min_margin"])
print("front end:", feats.shape, "decode tokens:", r["n_tokens"], "min margin:", r["min_margin"].min())
