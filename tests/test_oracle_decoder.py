"""Prediction net + joint oracle (PARITY UNPINNED against the absent decoder_joint ONNX; architecture pinned by its
byte size, SURVEY finding 2): the C loops are cross-checked against torch.nn.LSTM / Linear and frozen by a golden
fixture; the synthetic benchmark model's token rate is re-checked."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_parameter_count_matches_onnx_byte_size(oracle):
    n = 1025 * 640 + 2 * (2 * 2560 * 640 + 2 * 2560) + (640 * 1024 + 640) + (640 * 640 + 640) + (1030 * 640 + 1030)
    assert n == oracle.N_PARAMS == 8946310 and n * 4 == 35785240  # + 6819 B of graph = 35792059 B LFS object


def test_decoder_joint_against_torch(oracle):
    torch = pytest.importorskip("torch")
    m = oracle.Model(seed=11)
    t = {k: torch.from_numpy(v.copy()) for k, v in m.tensors().items()}
    lstm = torch.nn.LSTM(640, 640, num_layers=2)
    with torch.no_grad():
        for l in range(2):
            getattr(lstm, f"weight_ih_l{l}").copy_(t[f"w_ih{l}"])
            getattr(lstm, f"weight_hh_l{l}").copy_(t[f"w_hh{l}"])
            getattr(lstm, f"bias_ih_l{l}").copy_(t[f"b_ih{l}"])
            getattr(lstm, f"bias_hh_l{l}").copy_(t[f"b_hh{l}"])
    rng = np.random.default_rng(0)
    T, U = 3, 4
    enc = (0.5 * rng.standard_normal((1024, T))).astype(np.float32)
    tg = np.array([1024, 3, 77, 1000], np.int32)
    s1 = (0.1 * rng.standard_normal((2, 1, 640))).astype(np.float32)
    s2 = (0.1 * rng.standard_normal((2, 1, 640))).astype(np.float32)
    out, o1, o2 = m.decoder_joint(enc, tg, s1, s2)
    with torch.no_grad():
        x = t["emb"][torch.from_numpy(tg).long()].unsqueeze(1)          # [U, 1, 640]
        g, (h, c) = lstm(x, (torch.from_numpy(s1), torch.from_numpy(s2)))
        pred = g.squeeze(1) @ t["w_pred"].T + t["b_pred"]                  # [U, 640]
        e = torch.from_numpy(enc).T @ t["w_enc"].T + t["b_enc"]            # [T, 640]
        z = torch.tanh(e.unsqueeze(0) + pred.unsqueeze(1))                 # [U, T, 640]
        ref = z @ t["w_out"].T + t["b_out"]
    assert np.abs(out - ref.numpy()).max() < 2e-5
    assert np.abs(o1 - h.numpy()).max() < 2e-6 and np.abs(o2 - c.numpy()).max() < 2e-6
    assert np.all(m.tensors()["emb"][1024] == 0)  # blank row is the padding row


def test_golden_decode_fixture(oracle):
    import amira_b200 as A
    g = np.load(os.path.join(GOLD, "decode_golden.npz"))
    model = oracle.Model(blob=A.synthetic_weights(int(g["seed"])))
    for b in range(g["enc"].shape[0]):
        L = int(g["lens"][b])
        r = oracle.greedy_decode(np.ascontiguousarray(g["enc"][b, :, :L]), L, model)
        n = int(g["n_tokens"][b])
        assert r.tokens == g["tokens"][b, :n].tolist() and r.n_steps == int(g["n_steps"][b])


def test_library_random_init_equals_oracle_generator(oracle):
    import amira_b200 as A
    assert np.array_equal(A.random_weights(3456, 0.25), oracle.Model(seed=3456, blank_bias=0.25).blob)


def test_synthetic_model_token_rate(oracle):
    """bench.py's workload: the calibrated synthetic model emits ~0.1-0.6 tokens per encoder frame (not 0, not 30)."""
    import amira_b200 as A
    model = oracle.Model(blob=A.synthetic_weights(3456))
    rng = np.random.default_rng(2345)
    enc = (0.5 * rng.standard_normal((16, 1024, 40))).astype(np.float32)
    r = oracle.greedy_decode_batch(model, enc, threads=0)
    rate = r["n_tokens"].sum() / (16 * 40)
    assert 0.1 < rate < 0.7 and r["rc"] == 0
    assert len({int(x) for b in range(16) for x in r["tokens"][b, :r["n_tokens"][b]]}) > 10


def test_out_of_table_token_fails_like_onnx_gather(oracle):
    m = oracle.Model(seed=3456)
    blob = m.blob.copy()
    blob[-1030:][1027] += 100.0
    rng = np.random.default_rng(9)
    enc = (0.5 * rng.standard_normal((1024, 3))).astype(np.float32)
    r = oracle.greedy_decode(enc, 3, oracle.Model(blob=blob))
    assert r.rc == -1 and r.tokens == [1027]
