"""Debug: do chunked H2D + D2H copies on three streams overlap on this box?"""
import time, torch
n_enc, n_feat, n_pcm = 392_000_000, 394_000_000, 142_000_000
enc_h = torch.empty(n_enc, dtype=torch.float32).pin_memory()
feat_h = torch.empty(n_feat, dtype=torch.float32).pin_memory()
pcm_h = torch.empty(n_pcm, dtype=torch.float32).pin_memory()
enc_d = torch.empty(n_enc, dtype=torch.float32, device="cuda")
feat_d = torch.empty(n_feat, dtype=torch.float32, device="cuda")
pcm_d = torch.empty(n_pcm, dtype=torch.float32, device="cuda")
sA, sB, sC = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def chunks(n, k=8): return [(n * i // k, n * (i + 1) // k) for i in range(k)]
def run(do_enc, do_feat, do_pcm, order="enc_first"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    jobs = []
    if do_pcm: jobs.append((sC, pcm_d, pcm_h, n_pcm, True))
    if do_enc: jobs.append((sA, enc_d, enc_h, n_enc, True))
    if do_feat: jobs.append((sB, feat_d, feat_h, n_feat, False))
    for s, d, h, n, h2d in jobs:
        with torch.cuda.stream(s):
            for a, b in chunks(n):
                if h2d: d[a:b].copy_(h[a:b], non_blocking=True)
                else: h[a:b].copy_(d[a:b], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for name, args in (("enc H2D", (1, 0, 0)), ("feat D2H", (0, 1, 0)), ("pcm H2D", (0, 0, 1)), ("enc H2D + feat D2H", (1, 1, 0)),
                   ("enc H2D + pcm H2D", (1, 0, 1)), ("all three", (1, 1, 1))):
    run(*args); print(f"{name:22s} {min(run(*args) for _ in range(3)):7.2f} ms")
