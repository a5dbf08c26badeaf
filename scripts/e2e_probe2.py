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
t_stride = int((flens.max() + 31) // 32 * 32); T = int(elens.max())
ctx = A.Context(device_id=0); ctx.load_weights(A.synthetic_weights(3456))
ctx_fe = A.Context(device_id=0)
pcm_pin = torch.from_numpy(pcm).pin_memory()
enc_pin = torch.empty((B, 1024, T), dtype=torch.float32).pin_memory(); enc_pin.normal_(0, 0.5)
feats_pin = torch.empty((B, 128, t_stride), dtype=torch.float32).pin_memory()
tok_pin = torch.zeros((B, 200), dtype=torch.int32).pin_memory(); ntok_pin = torch.zeros(B, dtype=torch.int32).pin_memory()
flens_out = np.zeros(B, np.int64)
def fe(): ctx_fe.preprocess_pcm16_raw(pcm_pin.data_ptr(), offsets, B, feats_pin.data_ptr(), t_stride, flens_out)
def dec(): ctx.greedy_decode_raw(enc_pin.data_ptr(), B, T, elens, tok_pin.data_ptr(), ntok_pin.data_ptr(), None)
fe(); dec(); fe(); dec()
os.environ["AMIRA_DEBUG_TIMELINE"] = "1"
print("--- fe alone", file=sys.stderr); fe()
print("--- fe with decode", file=sys.stderr)
th = threading.Thread(target=fe); th.start(); dec(); th.join()
