import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import amira_b200 as A
from conftest import synth_pcm
from test_gpu_pipeline import VOCAB, _write_vocab, stub_encoder
_write_vocab()
ctx = A.Context(device_id=0)
ctx.load_weights(A.synthetic_weights(3456))
for gain in (2.0, 3.0, 4.0, 5.0, 6.0, 8.0):
    p = A.B200AsrPipeline(ctx, VOCAB, lambda f, g=gain: np.ascontiguousarray(g * stub_encoder(f), dtype=np.float32))
    n = [len(p.process_batch(synth_pcm(3.5, 900 + i).tobytes()).tokens) for i in range(6)]
    print(gain, n)
    p.close()
