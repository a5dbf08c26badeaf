"""The CPU oracle against every known-answer test the reference's own unit tests hold for the hot path
(SURVEY.md 8c), plus the KATs derived from the same fixtures.  Citations are relative to the reference root."""
import numpy as np


# ---- a1-a3: PCM conversion --------------------------------------------------------------------------------------
def test_bytes_to_f32_simd_equals_scalar(oracle):
    # src/asr/simd.rs:1294-1315 (test_bytes_to_f32_consistency): 1000 bytes i % 256, tolerance 1e-6 there; exact here
    data = bytes(i % 256 for i in range(1000))
    a, b = oracle.bytes_to_f32_samples(data), oracle.bytes_to_f32_simd(data)
    assert a.size == b.size == 500 and np.array_equal(a, b)
    assert np.array_equal(a, oracle.bytes_to_f32_optimized(data))  # even length: all three agree


def test_audio_processing_optimizations_kat(oracle):
    # src/performance_opts.rs:616-634: bytes [0,128,255,0,64,192] -> 3 samples; values derived (SURVEY 8c iv)
    out = oracle.bytes_to_f32_optimized(bytes([0, 128, 255, 0, 64, 192]))
    assert out.tolist() == [-1.0, 255 / 32768, -16320 / 32768]


def test_odd_trailing_byte_rule(oracle):
    # src/performance_opts.rs:26-30: (byte as i16) as f32 / 128.0 (zero-extended); src/asr/audio.rs:18-26 drops it
    assert oracle.bytes_to_f32_optimized(bytes([1, 0, 200])).tolist() == [1 / 32768, 200 / 128]
    assert oracle.bytes_to_f32_samples(bytes([1, 0, 200])).tolist() == [1 / 32768]
    assert oracle.bytes_to_f32_simd(bytes([1, 0, 200])).tolist() == [1 / 32768]
    assert oracle.bytes_to_f32_optimized(b"").size == 0


# ---- a7: frame gather -------------------------------------------------------------------------------------------
def test_tensor_view_frame_extraction(oracle):
    # src/asr/zero_copy.rs:255-272: data [1..6], shape [1,2,3], t = 1 -> [2, 5] (layout f*T + t)
    n, out = oracle.extract_frame_into(np.arange(1, 7, dtype=np.float32), [1, 2, 3], 1)
    assert n == 2 and out.tolist() == [2.0, 5.0]
    n, _ = oracle.extract_frame_into(np.arange(1, 7, dtype=np.float32), [1, 2, 3], 3)  # t out of range
    assert n == 0
    n, _ = oracle.extract_frame_into(np.arange(1, 7, dtype=np.float32), [1, 2, 3], 0, out_len=1)  # buffer too small
    assert n == 0


def test_transpose_map(oracle):
    # src/asr/simd.rs:1377-1391: [F=4, T=8] -> [T, F]; result[1] = 8, [4] = 1, [5] = 9
    F, T = 4, 8
    x = np.arange(F * T, dtype=np.float32)
    res = np.concatenate([oracle.extract_frame_into(x, [1, F, T], t)[1] for t in range(T)])
    assert (res[0], res[1], res[4], res[5]) == (0.0, 8.0, 1.0, 9.0)


# ---- a9: argmax -------------------------------------------------------------------------------------------------
def test_argmax_zero_copy(oracle):
    # src/asr/zero_copy.rs:274-281
    idx, val = oracle.argmax_zero_copy([0.1, 0.7, 0.2, 0.9, 0.3])
    assert idx == 3 and abs(val - 0.9) < 1e-7


def test_argmax_simd_kats(oracle):
    # src/asr/simd.rs:1418-1440
    assert oracle.argmax_zero_copy([1.0, 5.0, 3.0, 9.0, 2.0, 7.0, 4.0, 8.0, 6.0, 0.0]) == (3, 9.0)
    x = np.arange(1000, dtype=np.float32)
    x[500] = 2000.0
    assert oracle.argmax_zero_copy(x) == (500, 2000.0)


def test_argmax_tie_and_edge_rules(oracle):
    assert oracle.argmax_zero_copy([1.0, 1.0])[0] == 0          # first maximum wins (strict >)
    assert oracle.argmax_zero_copy([]) == (0, 0.0)              # empty -> (0, 0.0)
    assert oracle.argmax_zero_copy([0.0, np.nan, 1.0])[0] == 2  # NaN never wins unless at index 0
    assert oracle.argmax_zero_copy([np.nan, 5.0])[0] == 0


# ---- a6: the greedy loop with mock step functions ----------------------------------------------------------------
def _mock(logits, log=None):
    def step(frame, targets, s1, s2):
        if log is not None:
            log.append((frame.copy(), targets.copy(), s1.copy()))
        return np.asarray(logits, np.float32), s1 + 1.0, s2
    return step


def test_zero_copy_decoder_mock(oracle):
    # src/asr/decoder_optimized.rs:331-366: 2 features x 3 frames, mock logits [0.1, 0.2, 0.9] -> tokens non-empty.
    # Derived (SURVEY 8c i): exactly 30 step calls per frame -> 90 tokens, all 2.
    enc = np.array([1, 2, 3, 4, 5, 6], np.float32)
    log = []
    r = oracle.greedy_decode(enc, 3, step=_mock([0.1, 0.2, 0.9], log), single_step=False)
    assert r.rc == 0 and r.tokens == [2] * 90 and r.n_steps == 90 and r.frames_visited == 3
    assert log[0][0].tolist() == [1.0, 4.0] and log[30][0].tolist() == [2.0, 5.0]  # frame gather f*T + t
    # literal refeed: targets = [blank] ++ tokens of this call (zero_copy.rs:283-299 pins [1024, 1, 2, 3])
    assert log[0][1].tolist() == [1024] and log[3][1].tolist() == [1024, 2, 2, 2]
    # state replaced on every call
    assert r.states_1.ravel()[0] == 90.0


def test_max_total_tokens_cap(oracle):
    # derived (ii): T = 7 -> 200 tokens, stops inside frame 6 after 20 symbols
    enc = np.zeros(2 * 7, np.float32)
    r = oracle.greedy_decode(enc, 7, step=_mock([0.1, 0.2, 0.9]))
    assert len(r.tokens) == 200 and r.n_steps == 200 and r.frames_visited == 7


def test_blank_advances_and_state_is_carried(oracle):
    # derived (iii): blank-max mock -> 0 tokens, exactly T step calls, state updated T times (decoder_optimized.rs:154)
    logits = np.zeros(1030, np.float32)
    logits[1024] = 1.0
    r = oracle.greedy_decode(np.zeros(1024 * 5, np.float32), 5, step=_mock(logits))
    assert r.tokens == [] and r.n_steps == 5 and r.states_1.ravel()[0] == 5.0


def test_single_step_targets(oracle):
    # north_star mode: targets = [last emitted token of this call, else blank], U = 1
    seq = iter([7, 1024, 9, 9, 1024, 1024])
    log = []

    def step(frame, targets, s1, s2):
        log.append(targets.tolist())
        lg = np.zeros(1030, np.float32)
        lg[next(seq)] = 1.0
        return lg, s1, s2

    r = oracle.greedy_decode(np.zeros(1024 * 3, np.float32), 3, step=step, single_step=True)
    assert r.tokens == [7, 9, 9] and log == [[1024], [7], [7], [9], [9], [9]]


def test_decode_step_failure_aborts(oracle):
    # decoder_optimized.rs:148-152: any step error -> "Decode step failed"
    r = oracle.greedy_decode(np.zeros(1024 * 2, np.float32), 2, step=lambda *a: None)
    assert r.rc == -1 and r.tokens == []


def test_flat_argmax_over_whole_output(oracle):
    # src/triton/model.rs:713 + decoder_optimized.rs:163: argmax over ALL returned logits (U*1030 in refeed mode);
    # a maximum in row u >= 1 is pushed as the index u*1030 + v
    calls = []

    def step(frame, targets, s1, s2):
        calls.append(len(targets))
        lg = np.zeros(len(targets) * 1030, np.float32)
        if len(calls) == 1:
            lg[5] = 1.0               # first call: token 5
        else:
            lg[1030 + 1024] = 1.0     # second call: blank of row 1 wins -> index 2054, NOT a blank
        return lg, s1, s2

    r = oracle.greedy_decode(np.zeros(1024, np.float32), 1, step=step, single_step=False, max_symbols=2)
    assert r.tokens == [5, 2054] and calls == [1, 2]


def test_vocabulary_decode(oracle, tmp_path):
    # src/asr/types.rs:87-135
    p = tmp_path / "vocab.txt"
    p.write_text("<unk> 0\n▁the 1\ncat 2\n▁s at 3\nbroken line\n<blk> 1024\n", encoding="utf-8")
    v = oracle.Vocabulary.load_from_file(str(p))
    assert v.decode_tokens([1, 2, 999, 3]) == "thecat s at"   # unknown id skipped; leading space trimmed
    assert v.decode_tokens([]) == ""
