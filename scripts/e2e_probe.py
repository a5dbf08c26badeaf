"""Debug: wall-clock of the host-buffer C-ABI calls alone and concurrently (bench.py's e2e step)."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import amira_b200 as A
from bench import encoded_len, make_workload

B = 1024
pcm, offsets, lens = make_workload(B, 4567)
flens = lens // 160 + 1
elens = np.array([encoded_len(int(x)) for x in flens], np.int64)
t_stride = int((flens.max() + 31) // 32 * 32)
T = int(elens.max())
ctx = A.Context(device_id=0); ctx.load_weights(A.synthetic_weights(3456))
ctx_fe = A.Context(device_id=0)
pcm_pin = torch.from_numpy(pcm).pin_memory()
enc_pin = torch.empty((B, 1024, T), dtype=torch.float32).pin_memory(); enc_pin.normal_(0, 0.5)
feats_pin = torch.empty((B, 128, t_stride), dtype=torch.float32).pin_memory()
tok_pin = torch.zeros((B, 200), dtype=torch.int32).pin_memory()
ntok_pin = torch.zeros(B, dtype=torch.int32).pin_memory()
flens_out = np.zeros(B, np.int64)
dev_a = torch.empty(max(enc_pin.numel(), feats_pin.numel()), dtype=torch.float32, device="cuda")

def fe(): ctx_fe.preprocess_pcm16_raw(pcm_pin.data_ptr(), offsets, B, feats_pin.data_ptr(), t_stride, flens_out)
def dec(): ctx.greedy_decode_raw(enc_pin.data_ptr(), B, T, elens, tok_pin.data_ptr(), ntok_pin.data_ptr(), None)
def both():
    th = threading.Thread(target=fe); th.start(); dec(); th.join()
def h2d(): dev_a[:enc_pin.numel()].copy_(enc_pin.view(-1), non_blocking=True); torch.cuda.synchronize()
def d2h(): feats_pin.view(-1).copy_(dev_a[:feats_pin.numel()], non_blocking=True); torch.cuda.synchronize()
def duplex():
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1): dev_a[:enc_pin.numel()].copy_(enc_pin.view(-1), non_blocking=True)
    with torch.cuda.stream(s2): feats_pin.view(-1).copy_(dev_a[:feats_pin.numel()], non_blocking=True)
    torch.cuda.synchronize()
for name, fn, nbytes in (("fe host->host", fe, 0), ("decode host->host", dec, 0), ("both threads", both, 0),
                        ("raw H2D enc", h2d, enc_pin.numel() * 4), ("raw D2H feats", d2h, feats_pin.numel() * 4),
                        ("raw duplex", duplex, enc_pin.numel() * 4 + feats_pin.numel() * 4)):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name:20s} {dt * 1e3:8.2f} ms" + (f"  {nbytes / dt / 1e9:6.1f} GB/s" if nbytes else ""))

# who waits for whom in the concurrent run
for trial in range(2):
    marks = {}
    def fe_t():
        marks["fe0"] = time.perf_counter(); fe(); marks["fe1"] = time.perf_counter()
    t0 = time.perf_counter()
    th = threading.Thread(target=fe_t); th.start()
    marks["d0"] = time.perf_counter(); dec(); marks["d1"] = time.perf_counter()
    th.join()
    print({k: round((v - t0) * 1e3, 2) for k, v in marks.items()})
# decode first, FE delayed by 5 ms / FE first, decode delayed by 12 ms
for delay_fe, delay_dec in ((0.005, 0.0), (0.0, 0.012)):
    marks = {}
    def fe_t():
        time.sleep(delay_fe); marks["fe0"] = time.perf_counter(); fe(); marks["fe1"] = time.perf_counter()
    t0 = time.perf_counter()
    th = threading.Thread(target=fe_t); th.start()
    time.sleep(delay_dec); marks["d0"] = time.perf_counter(); dec(); marks["d1"] = time.perf_counter()
    th.join()
    print("delays", delay_fe, delay_dec, {k: round((v - t0) * 1e3, 2) for k, v in marks.items()})
