import sys, numpy as np
sys.path.insert(0, "/root/repo")
import amira_b200 as A
rng = np.random.default_rng(0)
with A.Context(device_id=0) as ctx:
    ctx.load_weights(A.synthetic_weights(3456))
    enc = (0.5 * rng.standard_normal((4, 1024, 12))).astype(np.float32)
    try:
        toks, _, steps = ctx.greedy_decode(enc)
        print("ok", [len(t) for t in toks])
    except Exception as e:
        print("ERR", e)
