"""The mel front-end oracle: PARITY UNPINNED against the reference (its preprocessor ONNX is an absent LFS object), so
the C restatement is cross-checked by independent implementations (numpy float64, torch.stft + torchaudio filterbank)
and frozen by a golden fixture (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import synth_pcm

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_features_len_rule(oracle):
    # verified against torch.stft(center=True) in SURVEY 8c: 2560 -> 17, 160000 -> 1001, 480000 -> 3001
    assert [oracle.features_len(n) for n in (0, 1, 159, 160, 2560, 160000, 480000)] == [0, 1, 1, 2, 17, 1001, 3001]


def test_filterbank_matches_torchaudio(oracle):
    ta = pytest.importorskip("torchaudio")
    fb = oracle.mel_filterbank()
    ref = ta.functional.melscale_fbanks(257, 0.0, 8000.0, 128, 16000, norm="slaney", mel_scale="slaney").numpy().T
    assert fb.shape == (128, 257) and np.abs(fb - ref).max() < 1e-6  # torchaudio builds it in float32
    assert np.count_nonzero(fb) == 504 and (np.count_nonzero(fb, axis=0) <= 2).all() and (np.count_nonzero(fb, axis=1) <= 12).all()
    assert np.array_equal(fb, oracle._slaney_fb_numpy())


def test_c_oracle_equals_numpy_float64(oracle):
    w = synth_pcm(1.7, 3).astype(np.float32) / 32768.0
    a, L = oracle.preprocess(w, "f64")
    b, L2 = oracle.preprocess_numpy(w)
    assert L == L2 == w.size // 160 + 1
    assert np.abs(a - b).max() < 2e-6


def test_c_oracle_against_torch_stft(oracle):
    torch = pytest.importorskip("torch")
    w = synth_pcm(2.0, 4).astype(np.float32) / 32768.0
    x = torch.from_numpy(w).double()
    y = torch.cat([x[:1], x[1:] - 0.97 * x[:-1]])
    win = torch.hann_window(400, periodic=False, dtype=torch.float32).double()
    spec = torch.stft(y, 512, hop_length=160, win_length=400, window=win, center=True, pad_mode="reflect", return_complex=True)
    power = spec.abs() ** 2
    mel = torch.from_numpy(oracle.mel_filterbank()).double() @ power
    lm = torch.log(mel + 2.0 ** -24)
    ref = ((lm - lm.mean(dim=1, keepdim=True)) / (lm.std(dim=1, keepdim=True) + 1e-5)).numpy()
    got, L = oracle.preprocess(w, "f64")
    assert ref.shape == (128, L) and np.abs(got - ref).max() < 2e-6


def test_f32_arithmetic_is_not_enough_for_the_contract(oracle):
    # why the CUDA kernel runs its FFT in fp64: a float32 restatement misses 1e-4 on cfg2-like signals
    w = synth_pcm(10.0, 1234).astype(np.float32) / 32768.0
    a, _ = oracle.preprocess(w, "f32")
    b, _ = oracle.preprocess(w, "f64")
    assert 1e-4 < np.abs(a - b).max() < 1e-2


def test_edge_lengths(oracle):
    for n in (1, 2, 159, 160, 161, 255, 256, 257, 513):
        w = synth_pcm(1.0, 7)[:n].astype(np.float32) / 32768.0
        a, L = oracle.preprocess(w, "f64", t_stride=8)
        b, _ = oracle.preprocess_numpy(w)
        assert L == n // 160 + 1 and np.all(a[:, L:] == 0)
        if L > 1:
            assert np.abs(a[:, :L] - b).max() < 1e-5


def test_golden_frontend_fixture(oracle):
    g = np.load(os.path.join(GOLD, "frontend_golden.npz"))
    feats, L = oracle.preprocess(g["pcm"].astype(np.float32) / 32768.0, "f64")
    assert L == int(g["features_len"])
    assert np.abs(feats - g["features"]).max() < 1e-6


def test_batch_helper(oracle):
    pcms = [synth_pcm(0.4, 1), synth_pcm(0.25, 2)]
    offs = np.array([0, pcms[0].size, pcms[0].size + pcms[1].size])
    feats, lens = oracle.preprocess_pcm16_batch(np.concatenate(pcms), offs, 48, threads=2)
    assert lens.tolist() == [41, 26]
    one, _ = oracle.preprocess(pcms[1].astype(np.float32) / 32768.0, "f32", t_stride=48)
    assert np.array_equal(feats[1], one)
