import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only; see oracle/amira_oracle.h)."""
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def amira():
    import amira_b200 as A
    A.load_library()  # raises when libamira_b200.so has not been built: no fallback
    return A


@pytest.fixture(scope="session")
def ctx(amira):
    c = amira.Context(device_id=0)
    yield c
    c.close()


def synth_pcm(seconds: float, seed: int) -> np.ndarray:
    """BASELINE config 2 signal: 0.1 N(0,1) + three 0.2-amplitude sines in [100, 4000] Hz, clipped, i16."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * 16000))
    t = np.arange(n)
    x = 0.1 * rng.standard_normal(n)
    for _ in range(3):
        x += 0.2 * np.sin(2 * np.pi * rng.uniform(100, 4000) * t / 16000)
    return np.round(np.clip(x, -1, 1) * 32767).astype(np.int16)


def calibrated_weights(O=None, seed=3456, blank_bias=None):
    """The synthetic benchmark model (amira_b200.synthetic_weights): seeded, rescaled, blank-calibrated."""
    import amira_b200 as A
    return A.synthetic_weights(seed) if blank_bias is None else A.synthetic_weights(seed, blank_bias)


def oracle_row_sigma(oracle, wave):
    """sigma over time of every un-normalised log-mel row, from the independent float64 numpy restatement."""
    x = wave.astype(np.float64)
    n = x.size
    if n < 160:  # fewer than two frames
        return np.zeros(128)
    y = np.empty_like(x)
    y[0] = x[0]
    y[1:] = x[1:] - 0.97 * x[:-1]
    idx = np.arange(-256, n + 256)
    p = 2 * (n - 1)
    idx = np.mod(idx, p)
    idx = np.where(idx < n, idx, p - idx)
    ypad = y[idx]
    L = n // 160 + 1
    win = oracle.hann_window_padded()
    frames = np.stack([ypad[t * 160:t * 160 + 512] for t in range(L)]) * win[None, :]
    power = np.abs(np.fft.rfft(frames, axis=1)) ** 2
    logmel = np.log(power @ oracle.mel_filterbank().astype(np.float64).T + 2.0 ** -24)
    return logmel.std(axis=0, ddof=1)


def feature_bound(sigma, tol=1e-4):
    """Per-row error bound of the normalised features: the contract's 1e-4, or two fp32 quanta of the log-mel intermediate over the
    row's own sigma where the row is (nearly) constant in time (tests/test_gpu_parity_baseline.py explains why)."""
    return np.maximum(tol, 4e-6 / (sigma + 1e-5))
