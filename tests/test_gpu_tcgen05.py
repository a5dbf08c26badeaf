"""tcgen05 split-bf16 GEMM (TMA -> swizzled smem -> tcgen05.mma -> TMEM -> tcgen05.ld) against float64 numpy.
Proves the UMMA shared-memory / instruction descriptor conventions the decode engine relies on."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 640), (300, 200, 1280), (1, 1030, 640), (257, 640, 1024)])
def test_split_bf16_gemm_matches_float64(ctx, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    got = ctx.debug_tc_gemm(A, W, b)
    ref = A.astype(np.float64) @ W.astype(np.float64).T + b
    scale = np.abs(A.astype(np.float64)) @ np.abs(W.astype(np.float64)).T + 1.0
    # hi+lo keeps 16 mantissa bits per operand: error per product <= ~2^-16, far below a single-bf16 GEMM (2^-8)
    assert np.max(np.abs(got - ref) / scale) < 3e-5
    assert np.max(np.abs(got - ref)) < 1e-3
