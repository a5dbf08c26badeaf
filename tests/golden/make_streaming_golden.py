"""Generates tests/golden/streaming_golden.json from the Python oracle (oracle/streaming.py): transcript-weaving,
window-sequence and overlap-silence cases.  The reference has no fixtures for these functions (src/asr/weaving.rs,
src/asr/audio.rs carry no tests), so these vectors pin the restatement, not the reference.
Run from the repo root:  python tests/golden/make_streaming_golden.py"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import streaming as S  # noqa: E402

WORDS = "the quick brown fox jumps over a lazy dog and then runs far away from here while it rains señor naïve ▁x 漢字".split()
rng = random.Random(20251018)
weave = []
for _ in range(60):
    words = [rng.choice(WORDS) for _ in range(rng.randint(3, 14))]
    cut = rng.randint(1, len(words) - 1)
    ov = rng.randint(0, min(4, cut))
    second = words[cut - ov:]
    if rng.random() < 0.3:
        second[0] = rng.choice(WORDS)
    first, second = " ".join(words[:cut]), " ".join(second)
    pct = rng.choice([0.2, 0.3, 3 / 7, 0.5, 0.8])
    o, s = S.best_alignment(first, second, np.float32(pct))
    weave.append({"first": first, "second": second, "pct": pct, "overlap": int(o), "score": float(s),
                  "woven": S.weave_transcript_segs(first, second, np.float32(pct))})
windows = []
for total, win, lead, trail in [(160000, 56000, 16000, 8000), (56001, 56000, 16000, 8000), (30000, 56000, 16000, 8000),
                                (123457, 40000, 8000, 4000), (100000, 48000, 0, 0), (1, 56000, 16000, 8000)]:
    w = S.window_sequence(total, win, lead, trail)
    windows.append({"args": [total, win, lead, trail], "slices": [[a[0], a[1], b[0], b[1]] for a, b, _ in w],
                    "overlap": [float(o) for _, _, o in w]})
nrng = np.random.default_rng(3)
silence = []
for n, scale, mean_amp in [(37, 0.02, 0.01), (800, 0.02, 0.04), (801, 0.3, 0.6), (2400, 1e-4, 0.0), (2400, 0.02, 0.16), (1, 0.5, 0.1)]:
    a = (scale * nrng.standard_normal(n)).astype(np.float32)
    silence.append({"audio": [float(x) for x in a], "mean_amplitude": mean_amp, "mean_abs": float(S.mean_amplitude(a)),
                    "silent": bool(S.is_overlap_silence(a, np.float32(mean_amp)))})
with open(os.path.join(HERE, "streaming_golden.json"), "w", encoding="utf-8") as f:
    json.dump({"weave": weave, "windows": windows, "silence": silence}, f, ensure_ascii=False, indent=0)
print(len(weave), "weave cases,", sum(c["overlap"] > 0 and c["woven"] != c["first"] + " " + c["second"] for c in weave), "aligned")
