"""Soak of the tcgen05 decode kernel (debug): the bench workload's decode repeated under different lane plans and speculation
tables (AMIRA_WS_TILES / AMIRA_WS_SPEC / AMIRA_WS_PAIR change WHEN every step is computed, never its inputs): every run must give the
same tokens, counts and step counts as the first.  Usage: python scripts/decode_soak.py [runs per variant]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = [{}, {"AMIRA_WS_TILES": "2"}, {"AMIRA_WS_TILES": "3"}, {"AMIRA_WS_TILES": "8", "AMIRA_WS_SPEC": "0"}, {"AMIRA_WS_SPEC": "1,1,1,1,1,1,1,1"},
            {"AMIRA_WS_SPEC": "3,3,3,3"}, {"AMIRA_WS_PAIR": "0"}, {"AMIRA_WS_CHUNK": "0"}, {"AMIRA_WS_TILES": "6", "AMIRA_WS_SPEC": "2,2,2,2,2,2"}]


def child(runs):
    import torch
    import amira_b200 as A
    from bench import encoded_len, make_workload
    B = 1024
    ctx = A.Context(device_id=0, decode_engine=4)
    ctx.load_weights(A.synthetic_weights(3456))
    _, _, lens = make_workload(B, 4567)
    elens = np.array([encoded_len(int(x // 160 + 1)) for x in lens], np.int64)
    T = int(elens.max())
    g = torch.Generator(device="cuda")
    g.manual_seed(2345)
    enc = torch.randn((B, 1024, T), generator=g, device="cuda", dtype=torch.float32) * 0.5
    out = []
    for _ in range(runs):
        tok = torch.zeros((B, 200), dtype=torch.int32, device="cuda")
        nt = torch.zeros(B, dtype=torch.int32, device="cuda")
        ns = torch.zeros(B, dtype=torch.int32, device="cuda")
        ctx.greedy_decode_raw(enc.data_ptr(), B, T, elens, tok.data_ptr(), nt.data_ptr(), ns.data_ptr())
        torch.cuda.synchronize()
        n = nt.cpu().numpy()
        t = tok.cpu().numpy()
        out.append((n.copy(), ns.cpu().numpy().copy(), np.concatenate([t[b, :max(n[b], 0)] for b in range(B)])))
    np.savez(sys.argv[3], n=np.stack([o[0] for o in out]), s=np.stack([o[1] for o in out]), t=np.stack([o[2] for o in out]))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
        sys.exit(0)
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    ref = None
    bad = 0
    for i, v in enumerate(VARIANTS):
        path = f"/tmp/decode_soak_{i}.npz"
        env = dict(os.environ, **v)
        subprocess.run([sys.executable, __file__, "--child", str(runs), path], check=True, env=env)
        d = np.load(path)
        if ref is None:
            ref = (d["n"][0], d["s"][0], d["t"][0])
        ok = all(np.array_equal(d["n"][r], ref[0]) and np.array_equal(d["s"][r], ref[1]) and np.array_equal(d["t"][r], ref[2]) for r in range(runs))
        bad += 0 if ok else 1
        print(f"variant {v or 'default'}: {runs} runs, tokens {int(d['n'][0].clip(0).sum())}, {'identical' if ok else 'MISMATCH'}", flush=True)
    print("soak", "FAILED" if bad else "ok")
    sys.exit(1 if bad else 0)
