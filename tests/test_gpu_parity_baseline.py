"""GPU parity against the CPU ORACLE at BASELINE.json sizes (not GPU-vs-GPU properties):

  cfg2  preprocessor, 64 x 30 s      4 distinct 30 s utterances replicated to 64, every row vs the f64 oracle
  cfg3  greedy loop, 256 streams     8 distinct streams x T = 126 replicated to 256 (oracle on the 8); also T = 376
  cfg5  mixed 5-30 s shard           128 streams with the encoded lengths bench.py draws (T <= 375), tokens, steps and
                                     final DecoderState vs the oracle

Token IDs are exact; a stream may differ from the oracle only when the oracle's own top-1/top-2 margin somewhere in that
stream is below NEAR_TIE (documented exception, BASELINE.json north_star).  The recurrent state is compared where the
tokens agree: the split-bf16 error compounds over up to ~480 sequential steps here, which is the regime the small tests
(T <= 40) do not reach.  The bound is CALIBRATED: with the synthetic weights the cell state grows to |c| ~ 100-400, so ANY
fp32 implementation drifts from exact arithmetic by ~1e-3 there (the fp32 oracle is 3e-3 off a float64 run at T = 376);
the GPU must be no further from the float64 run (tests/f64_decoder.py) than 3x the fp32 oracle's own distance + 1e-4.
Reference: src/asr/decoder_optimized.rs:54-200, src/constants.rs:133-137."""
import os
import sys

import numpy as np
import pytest

from conftest import feature_bound, oracle_row_sigma, synth_pcm
from f64_decoder import greedy_decode_f64

pytestmark = pytest.mark.gpu
NEAR_TIE = 2e-4
STATE_ABS, STATE_REL = 1e-4, 3.0  # |gpu - f64| <= STATE_REL * |fp32 oracle - f64| + STATE_ABS
FEAT_TOL = 1e-4


def _encoded_len(L: int) -> int:
    for _ in range(3):
        L = (L - 1) // 2 + 1 if L > 0 else 0
    return L


@pytest.fixture(scope="module")
def dctx(amira):
    c = amira.Context(device_id=0)
    c.load_weights(amira.synthetic_weights(3456))
    yield c
    c.close()


@pytest.fixture(scope="module")
def model(oracle, amira):
    return oracle.Model(blob=amira.synthetic_weights(3456))


def _compare_with_oracle(oracle, model, enc, lens, toks, steps, st, n_f64=8):
    """Tokens / steps of every stream against the fp32 oracle; final state of the first n_f64 streams against a float64 run,
    calibrated by the fp32 oracle's own distance from it.  Returns (exact, near-tie, worst gpu-f64 err, worst oracle-f64 err)."""
    B = enc.shape[0]
    ref = oracle.greedy_decode_batch(model, enc, enc_lens=lens, states=(np.zeros((2, B, 640), np.float32), np.zeros((2, B, 640), np.float32)))
    assert ref["rc"] == 0
    n_exact = n_tie = 0
    exact = []
    for b in range(B):
        want = ref["tokens"][b, :int(ref["n_tokens"][b])].tolist()
        if toks[b] == want:
            assert int(steps[b]) == int(ref["n_steps"][b]), b
            n_exact += 1
            exact.append(b)
        else:
            assert float(ref["min_margin"][b]) < NEAR_TIE, (b, float(ref["min_margin"][b]), len(want), len(toks[b]))
            n_tie += 1
    n = min(n_f64, B)
    f_toks, f_steps, f_h, f_c, _ = greedy_decode_f64(model.blob, enc[:n], lens[:n])
    worst_gpu = worst_orc = 0.0
    for b in [b for b in exact if b < n]:
        if f_toks[b] != toks[b]:
            continue  # the float64 run itself took another branch at a near-tie: nothing to calibrate against
        e_gpu = max(np.abs(st.states_1[:, b] - f_h[:, b]).max(), np.abs(st.states_2[:, b] - f_c[:, b]).max())
        e_orc = max(np.abs(ref["states_1"][:, b] - f_h[:, b]).max(), np.abs(ref["states_2"][:, b] - f_c[:, b]).max())
        assert e_gpu <= STATE_REL * e_orc + STATE_ABS, (b, float(e_gpu), float(e_orc))
        worst_gpu, worst_orc = max(worst_gpu, float(e_gpu)), max(worst_orc, float(e_orc))
    return n_exact, n_tie, worst_gpu, worst_orc


@pytest.mark.parametrize("T", [126, 376])
def test_cfg3_256_streams_equal_oracle(dctx, oracle, model, T):
    """BASELINE config 3: 256 streams (8 distinct, replicated), limits 30/200, zero initial state."""
    rng = np.random.default_rng(2345)
    base = (0.5 * rng.standard_normal((8, 1024, T))).astype(np.float32)
    B = 256
    enc = np.ascontiguousarray(base[np.arange(B) % 8])
    lens = np.full(B, T, np.int64)
    toks, st, steps = dctx.greedy_decode(enc, lens)
    n_exact, n_tie, e_gpu, e_orc = _compare_with_oracle(oracle, model, base, lens[:8], toks[:8], steps[:8], _first(st, 8))
    print(f"cfg3 T={T}: exact {n_exact}/8, near-tie {n_tie}, final state vs float64: gpu {e_gpu:.2e}, fp32 oracle {e_orc:.2e}, "
          f"steps {steps[:8].tolist()}")
    assert n_tie <= 1
    for b in range(8, B):  # replicas decode exactly like their originals (batch position must not matter)
        assert toks[b] == toks[b % 8] and steps[b] == steps[b % 8], b
        assert np.array_equal(st.states_1[:, b], st.states_1[:, b % 8])


def _first(st, n):
    import amira_b200 as A
    return A.DecoderState(np.ascontiguousarray(st.states_1[:, :n]), np.ascontiguousarray(st.states_2[:, :n]))


def test_cfg5_128_streams_bench_lengths_equal_oracle(dctx, oracle, model):
    """BASELINE config 5 shard: the first 128 utterance lengths bench.py draws (U[5,30] s, seed 4567 => T in 63..375),
    packed ragged encoder outputs, tokens + steps + final state vs the oracle."""
    secs = np.random.default_rng(4567).uniform(5.0, 30.0, size=1024)[:128]
    elens = np.array([_encoded_len(int(round(s * 16000)) // 160 + 1) for s in secs], np.int64)
    T = int(elens.max())
    rng = np.random.default_rng(2345)
    enc = (0.5 * rng.standard_normal((128, 1024, T))).astype(np.float32)
    ragged = [np.ascontiguousarray(enc[b, :, :int(elens[b])]) for b in range(128)]
    toks, st, steps = dctx.greedy_decode_packed(ragged)
    for b in range(128):
        enc[b, :, int(elens[b]):] = 0.0
    n_exact, n_tie, e_gpu, e_orc = _compare_with_oracle(oracle, model, enc, elens, toks, steps, st, n_f64=16)
    ntok = sum(len(t) for t in toks)
    print(f"cfg5: exact {n_exact}/128, near-tie {n_tie}, final state vs float64 (16 streams): gpu {e_gpu:.2e}, fp32 oracle {e_orc:.2e}, "
          f"sum steps {int(steps.sum())}, tokens {ntok}, T {int(elens.min())}..{T}")
    assert n_tie <= 3
    assert int(steps.max()) > 400  # the long-chain regime is really exercised


def test_cfg2_64x30s_equal_f64_oracle(dctx, oracle):
    """BASELINE config 2: 64 x 30 s.  4 distinct utterances (oracle f64 on each), replicated to 64 rows."""
    base = [synth_pcm(30.0, 1234 + i) for i in range(4)]
    pcms = [base[i % 4] for i in range(64)]
    offs = np.zeros(65, np.int64)
    offs[1:] = np.cumsum([p.size for p in pcms])
    feats, lens = dctx.preprocess_pcm16(np.concatenate(pcms), offs, t_stride=3008)
    assert np.all(lens == 3001)
    worst = 0.0
    for i in range(4):
        ref, L = oracle.preprocess(base[i].astype(np.float32) / 32768.0, "f64")
        assert L == 3001
        worst = max(worst, float(np.abs(feats[i, :, :L] - ref).max()))
    print(f"cfg2: max |features - f64 oracle| = {worst:.2e}")
    assert worst <= FEAT_TOL
    assert np.all(feats[:, :, 3001:] == 0)
    for i in range(4, 64):
        assert np.array_equal(feats[i], feats[i % 4]), i


# ---------------------------------------------------------------------------------------------- adversarial front-end signals
def _adversarial_signals():
    n = 32000  # 2 s
    t = np.arange(n)
    rng = np.random.default_rng(77)
    sq = np.where((t // 16) % 2 == 0, 32767, -32768).astype(np.int16)
    sigs = {
        "digital_silence": np.zeros(n, np.int16),
        "dc_1000": np.full(n, 1000, np.int16),
        "dc_full_scale": np.full(n, 32767, np.int16),
        "dc_negative_full_scale": np.full(n, -32768, np.int16),
        "square_full_scale_500hz": sq,
        "square_clipped_noise": np.clip(sq.astype(np.int32) + rng.integers(-3, 4, n), -32768, 32767).astype(np.int16),
        "tone_1khz_full_scale": np.round(32767 * np.sin(2 * np.pi * 1000.0 * t / 16000)).astype(np.int16),
        "tone_440hz_full_scale": np.round(32767 * np.sin(2 * np.pi * 440.0 * t / 16000)).astype(np.int16),
        "tone_440hz_quiet": np.round(40 * np.sin(2 * np.pi * 440.0 * t / 16000)).astype(np.int16),
        "tone_7900hz": np.round(20000 * np.sin(2 * np.pi * 7900.0 * t / 16000)).astype(np.int16),
        "one_lsb_noise": rng.integers(-1, 2, n).astype(np.int16),
        "impulse": np.concatenate([np.zeros(12345, np.int16), np.array([32767], np.int16), np.zeros(n - 12346, np.int16)]),
        "silence_then_noise": np.concatenate([np.zeros(n // 2, np.int16), (3000 * rng.standard_normal(n // 2)).astype(np.int16)]),
        "alternating_full_scale": np.where(t % 2 == 0, 32767, -32768).astype(np.int16),
    }
    return sigs


def test_adversarial_signals_equal_f64_oracle(dctx, oracle):
    """Signals where (x - mu) / (sigma + 1e-5) and log(P + 2^-24) amplify rounding: digital silence, DC, full-scale squares with
    clipping, pure tones (periodic in the hop and not), 1-LSB noise, a single impulse.  Bound per (utterance, mel) row with
    sigma = the row's standard deviation over time in the float64 restatement:

        |err| <= max(1e-4, 4e-6 / (sigma + 1e-5))

    i.e. the contract's 1e-4 wherever a row is not numerically constant; for (near-)constant rows the quotient is rounding noise
    over ~1e-5 in ANY fp32 implementation (the fp32 oracle itself is 1e-2 off the float64 one on these rows), so they are
    bounded in un-normalised log-mel units instead (4e-6 = two fp32 quanta of the log-mel intermediate at |log 2^-24| = 16.6)."""
    sigs = _adversarial_signals()
    names = list(sigs)
    pcms = [sigs[k] for k in names]
    offs = np.zeros(len(pcms) + 1, np.int64)
    offs[1:] = np.cumsum([p.size for p in pcms])
    feats, lens = dctx.preprocess_pcm16(np.concatenate(pcms), offs)
    report = {}
    for b, k in enumerate(names):
        w = pcms[b].astype(np.float32) / 32768.0
        ref, L = oracle.preprocess(w, "f64")
        assert int(lens[b]) == L
        got = feats[b, :, :L]
        assert np.all(np.isfinite(got)), k
        sigma = oracle_row_sigma(oracle, w)
        err = np.abs(got - ref).max(axis=1)
        bound = feature_bound(sigma, FEAT_TOL)
        regular = sigma >= 2e-2
        report[k] = (float(err[regular].max()) if regular.any() else 0.0, float((err / bound).max()), int((~regular).sum()))
        assert np.all(err <= bound), (k, int(np.argmax(err / bound)), float((err / bound).max()))
    assert np.abs(feats[names.index("digital_silence")]).max() <= 1e-6  # constant log-mel: 0 in exact arithmetic
    print("adversarial front-end signals: (max err on rows with sigma >= 2e-2, max err / bound over all rows, rows below 2e-2)")
    for k, v in report.items():
        print(f"  {k:28s} {v[0]:.2e} {v[1]:.2f} {v[2]}")


def test_one_frame_utterance_is_exactly_zero(dctx):
    """features_len = 1: sigma = 0 and x - mean = 0 exactly, so the contract value is 0 (the oracle gives 0 too)."""
    pcm = synth_pcm(0.005, 3)  # 80 samples -> one frame
    feats, lens = dctx.preprocess_pcm16(pcm, [0, pcm.size])
    assert int(lens[0]) == 1 and np.all(feats[0, :, 0] == 0.0)


# ---------------------------------------------------------------------------------------------- a9: argmax tie / NaN rule on the GPU
def _tie_blob(amira, j1, j2, bias=60.0):
    blob = amira.synthetic_weights(3456)
    t = amira.blob_views(blob)
    t["w_out"][j2] = t["w_out"][j1]
    t["b_out"][j1] = t["b_out"][j2] = np.float32(bias)
    return blob


@pytest.mark.parametrize("engine", [1, 4])
@pytest.mark.parametrize("pair", [(3, 17), (5, 40), (10, 700), (100, 1024), (0, 1023), (63, 64)])
def test_exact_tie_lower_index_wins(amira, oracle, engine, pair):
    """src/asr/zero_copy.rs:190-232: strict '>' from seed index 0, so among exactly equal maxima the LOWEST index wins.  Two
    vocabulary rows are made bit-identical (same weights, same bias, large enough to be the maximum at every step): the pairs
    sit in one thread's columns, in the two column groups of one 64-wide slice, in different slices (different SMs, merged by
    the 64-bit atomicMax key) and across the blank id.  AMIRA_WS_NOROT makes every slice accumulate its k-chunks in the same
    order, so the two logits are bit-identical across SMs as well."""
    j1, j2 = pair
    blob = _tie_blob(amira, j1, j2)
    rng = np.random.default_rng(9)
    enc = (0.5 * rng.standard_normal((3, 1024, 4))).astype(np.float32)
    r = oracle.greedy_decode(enc[0], 4, oracle.Model(blob=blob))
    assert r.tokens == [j1] * 120  # oracle: the lower index at every step, 30 per frame
    os.environ["AMIRA_WS_NOROT"] = "1"
    try:
        with amira.Context(device_id=0, decode_engine=engine) as c:
            c.load_weights(blob)
            toks, _, steps = c.greedy_decode(enc, [4, 4, 2])
    finally:
        os.environ.pop("AMIRA_WS_NOROT", None)
    assert toks[0] == [j1] * 120 and toks[1] == [j1] * 120 and toks[2] == [j1] * 60
    assert steps.tolist() == [120, 120, 60]


@pytest.mark.parametrize("engine", [1, 4])
@pytest.mark.parametrize("col", [5, 32, 64, 640, 1024])
def test_nan_logit_never_wins_unless_first(amira, oracle, engine, col):
    """zero_copy.rs:190-232: `logits[i] > max` is false for NaN, so a NaN logit can only be returned when it is the seed
    (index 0).  A NaN bias at any other column — including the first column of a thread's / a slice's range — must leave the
    decode identical to the one with that column removed from the competition."""
    blob = amira.synthetic_weights(3456)
    amira.blob_views(blob)["b_out"][col] = np.float32(np.nan)
    rng = np.random.default_rng(2345)
    enc = (0.5 * rng.standard_normal((2, 1024, 40))).astype(np.float32)
    model = oracle.Model(blob=blob)
    with amira.Context(device_id=0, decode_engine=engine) as c:
        c.load_weights(blob)
        toks, _, steps = c.greedy_decode(enc)
    assert sum(len(t) for t in toks) > 0
    for b in range(2):
        r = oracle.greedy_decode(enc[b], 40, model)
        assert col not in r.tokens
        assert toks[b] == r.tokens or r.margins.min() < NEAR_TIE
        if col == 1024:
            assert len(toks[b]) > 0  # blank can never win: every step emits


@pytest.mark.parametrize("engine", [1, 4])
def test_nan_at_index_zero_always_wins(amira, oracle, engine):
    blob = amira.synthetic_weights(3456)
    amira.blob_views(blob)["b_out"][0] = np.float32(np.nan)
    rng = np.random.default_rng(11)
    enc = (0.5 * rng.standard_normal((1, 1024, 9))).astype(np.float32)
    r = oracle.greedy_decode(enc[0], 9, oracle.Model(blob=blob))
    assert r.tokens == [0] * 200
    with amira.Context(device_id=0, decode_engine=engine) as c:
        c.load_weights(blob)
        toks, _, steps = c.greedy_decode(enc)
    assert toks[0] == [0] * 200 and steps[0] == 200
