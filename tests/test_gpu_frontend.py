"""GPU parity of the fused front end against the CPU oracle (float64 restatement of the preprocessor spec).
Tolerance from BASELINE.json north_star: max abs error <= 1e-4 after normalisation."""
import numpy as np
import pytest

from conftest import feature_bound, oracle_row_sigma, synth_pcm

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _oracle_feats(O, pcm_list, t_stride):
    out = np.zeros((len(pcm_list), 128, t_stride), np.float32)
    lens = []
    for b, p in enumerate(pcm_list):
        w = p.astype(np.float32) / 32768.0
        if w.size:
            f, L = O.preprocess(w, "f64", t_stride=t_stride)
            out[b] = f
        else:
            L = 0
        lens.append(L)
    return out, np.array(lens)


def _pack(pcm_list):
    offs = np.zeros(len(pcm_list) + 1, np.int64)
    offs[1:] = np.cumsum([p.size for p in pcm_list])
    return (np.concatenate(pcm_list) if offs[-1] else np.zeros(0, np.int16)), offs


def test_bytes_to_f32_matches_reference_rule(ctx, oracle):
    raw = bytes((i * 37 + 11) % 256 for i in range(1001))  # odd length: trailing-byte rule, performance_opts.rs:26-30
    for data in (raw, raw[:-1], raw[:6], raw[:1], b"", bytes([0, 128, 255, 0, 64, 192])):
        got = ctx.bytes_to_f32(data)
        assert np.array_equal(got, oracle.bytes_to_f32_optimized(data))
        got = ctx.bytes_to_f32(data, drop_odd=True)
        assert np.array_equal(got, oracle.bytes_to_f32_samples(data))


def test_single_utterance_10s(ctx, oracle):
    pcm = synth_pcm(10.0, 1234)
    feats, lens = ctx.preprocess_pcm16(pcm, [0, pcm.size])
    ref, rl = _oracle_feats(oracle, [pcm], feats.shape[2])
    assert lens.tolist() == rl.tolist() == [1001]
    assert np.abs(feats - ref).max() <= TOL


def test_ragged_batch_with_padding_and_edge_lengths(ctx, oracle):
    secs = [0.5, 3.21, 1.0, 0.01, 2.0]
    pcms = [synth_pcm(s, 100 + i) for i, s in enumerate(secs)]
    pcms += [synth_pcm(1.0, 7)[:n] for n in (1, 2, 159, 160, 161, 255, 256, 257, 300, 511, 513, 5120, 5121)]
    pcms.insert(3, np.zeros(0, np.int16))  # empty utterance: features_len 0, all-zero row
    pcm, offs = _pack(pcms)
    t_stride = 328
    feats, lens = ctx.preprocess_pcm16(pcm, offs, t_stride=t_stride)
    ref, rl = _oracle_feats(oracle, pcms, t_stride)
    assert lens.tolist() == rl.tolist()
    for b in range(len(pcms)):
        L = int(lens[b])
        assert np.all(feats[b, :, L:] == 0.0), b
        if L > 1:
            # a few frames that share most of their (reflected) samples give nearly constant rows: per-row bound (conftest)
            sigma = oracle_row_sigma(oracle, pcms[b].astype(np.float32) / 32768.0)
            err = np.abs(feats[b] - ref[b]).max(axis=1)
            assert np.all(err <= feature_bound(sigma, TOL)), (b, pcms[b].size, float((err / feature_bound(sigma, TOL)).max()))
        elif L == 1:
            assert np.all(feats[b, :, 0] == 0.0)  # one frame: x - mean = 0 exactly


def test_f32_contract_form_matches_pcm_form(ctx, oracle):
    pcms = [synth_pcm(2.0, 5), synth_pcm(1.3, 6)]
    N = max(p.size for p in pcms)
    wav = np.zeros((2, N), np.float32)
    for b, p in enumerate(pcms):
        wav[b, :p.size] = p.astype(np.float32) / 32768.0
    feats, lens = ctx.preprocessor(wav, [p.size for p in pcms])
    ref, rl = _oracle_feats(oracle, pcms, feats.shape[2])
    assert lens.tolist() == rl.tolist()
    assert np.abs(feats - ref).max() <= TOL
    pcm, offs = _pack(pcms)
    f2, _ = ctx.preprocess_pcm16(pcm, offs, t_stride=feats.shape[2])
    assert np.abs(f2 - feats).max() <= 1e-5


def test_full_size_properties_64x30s(ctx):
    """BASELINE config 2 (64 x 30 s): size-independent properties — zero mean / unit variance per (utterance, mel)
    over valid frames, zero padding, batch invariance (row b equals the same utterance run alone)."""
    base = [synth_pcm(30.0, 1234 + i) for i in range(4)]
    pcms = [base[i % 4] for i in range(64)]
    pcm, offs = _pack(pcms)
    feats, lens = ctx.preprocess_pcm16(pcm, offs, t_stride=3008)
    assert np.all(lens == 3001)
    v = feats[:, :, :3001].astype(np.float64)
    assert np.abs(v.mean(axis=2)).max() < 1e-4
    assert np.abs(v.std(axis=2, ddof=1) - 1.0).max() < 1e-3
    assert np.all(feats[:, :, 3001:] == 0)
    for i in range(4, 64):
        assert np.array_equal(feats[i], feats[i % 4])
    alone, _ = ctx.preprocess_pcm16(base[1], [0, base[1].size], t_stride=3008)
    assert np.array_equal(alone[0], feats[1])


def test_invalid_arguments(ctx, amira):
    pcm = synth_pcm(1.0, 1)
    with pytest.raises(amira.AmiraError) as e:
        ctx.preprocess_pcm16(pcm, [0, pcm.size], t_stride=50)  # features_len is 101
    assert e.value.code == 1


def test_golden_fixture(ctx):
    """Committed golden vector (tests/golden/make_golden.py, produced by the float64 oracle)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "frontend_golden.npz"))
    feats, lens = ctx.preprocess_pcm16(g["pcm"], [0, g["pcm"].size])
    assert int(lens[0]) == int(g["features_len"])
    assert np.abs(feats[0, :, :int(lens[0])] - g["features"]).max() <= TOL


def test_pipelined_host_path_equals_resident_path(ctx, amira):
    """Host buffers take the chunked H2D / kernel / D2H pipeline (>= 64 utterances => several chunks); device buffers take
    the single-launch path.  Both must produce bit-identical features (utterances are independent)."""
    import torch
    rng = np.random.default_rng(11)
    pcms = [synth_pcm(float(rng.uniform(0.3, 1.2)), 300 + i) for i in range(96)]
    pcm, offs = _pack(pcms)
    t_stride = 128
    host, lens_h = ctx.preprocess_pcm16(pcm, offs, t_stride=t_stride)
    pcm_d = torch.from_numpy(pcm).cuda()
    out_d = torch.empty((len(pcms), 128, t_stride), dtype=torch.float32, device="cuda")
    lens_d = np.zeros(len(pcms), np.int64)
    ctx.preprocess_pcm16_raw(pcm_d.data_ptr(), offs, len(pcms), out_d.data_ptr(), t_stride, lens_d)
    torch.cuda.synchronize()
    assert lens_h.tolist() == lens_d.tolist()
    assert np.array_equal(host, out_d.cpu().numpy())


def test_packed_output_equals_padded_output(ctx, amira):
    """amira_preprocess_pcm16_packed: ragged [128][features_len_b] blocks (no padding to the longest utterance) must be
    bit-identical to the padded layout — through the chunked host pipeline (96 utterances), through device pointers,
    with an empty utterance, a one-frame utterance and a non-zero first offset."""
    import torch
    rng = np.random.default_rng(12)
    pcms = [synth_pcm(float(rng.uniform(0.2, 1.5)), 500 + i) for i in range(96)]
    pcms[3] = np.zeros(0, np.int16)
    pcms[7] = synth_pcm(0.005, 1)  # 80 samples -> one frame
    pcm, offs = _pack(pcms)
    padded, lens_p = ctx.preprocess_pcm16(pcm, offs, t_stride=160)
    blocks, lens_k = ctx.preprocess_pcm16_packed(pcm, offs)
    assert lens_p.tolist() == lens_k.tolist()
    for b in range(len(pcms)):
        assert blocks[b].shape == (128, int(lens_p[b]))
        assert np.array_equal(blocks[b], padded[b, :, :int(lens_p[b])]), b
    # device pointers, blocks with gaps and a non-zero first offset
    foff = np.zeros(len(pcms) + 1, np.int64)
    foff[0] = 5
    for b in range(len(pcms)):
        foff[b + 1] = foff[b] + 128 * int(lens_p[b]) + (b % 3)
    pcm_d = torch.from_numpy(pcm).cuda()
    out_d = torch.zeros(int(foff[-1]), dtype=torch.float32, device="cuda")
    lens_d = np.zeros(len(pcms), np.int64)
    ctx.preprocess_pcm16_packed_raw(pcm_d.data_ptr(), offs, len(pcms), out_d.data_ptr(), foff, lens_d)
    torch.cuda.synchronize()
    out = out_d.cpu().numpy()
    for b in range(len(pcms)):
        L = int(lens_p[b])
        assert np.array_equal(out[foff[b]:foff[b] + 128 * L].reshape(128, L), padded[b, :, :L]), b
    # a block that is too small is refused
    bad = foff.copy()
    bad[1:] -= 64
    with pytest.raises(amira.AmiraError):
        ctx.preprocess_pcm16_packed_raw(pcm_d.data_ptr(), offs, len(pcms), out_d.data_ptr(), bad, lens_d)
