"""Host logic of the streaming orchestrator: the C++ restatements behind the C ABI (csrc/host_stream.cpp)
against the Python oracle (oracle/streaming.py) — two independent restatements of src/asr/weaving.rs, src/asr/audio.rs and
src/performance_opts.rs:35-60.  The reference has no tests for these functions, so besides the cross-check the cases below pin
the properties the reference's code implies (identity on disjoint text, removal of a repeated overlap, window tiling).
No GPU is needed: these entries are pure host functions."""
import random

import numpy as np
import pytest

import streaming as S  # oracle/streaming.py


@pytest.fixture(scope="module")
def st(amira):
    return amira.streaming


WORDS = "the quick brown fox jumps over a lazy dog and then runs far away from here while it rains señor naïve ▁x 漢字".split()


def _sentence(rng, n):
    return " ".join(rng.choice(WORDS) for _ in range(n))


def test_weave_matches_oracle_on_random_overlapping_transcripts(st):
    rng = random.Random(5)
    n_woven = 0
    for case in range(120):
        words = [rng.choice(WORDS) for _ in range(rng.randint(3, 14))]
        cut = rng.randint(1, len(words) - 1)
        ov = rng.randint(0, min(4, cut))
        first = " ".join(words[:cut])
        second_words = words[cut - ov:]
        if rng.random() < 0.3 and second_words:  # a recognition error inside the overlap
            second_words[0] = rng.choice(WORDS)
        second = " ".join(second_words)
        pct = np.float32(rng.choice([0.2, 0.3, 3 / 7, 0.5, 0.8]))
        o_c, s_c = st.best_alignment(first, second, float(pct))
        o_o, s_o = S.best_alignment(first, second, pct)
        assert abs(s_c - float(s_o)) <= 1e-5 * max(1.0, abs(float(s_o))), (first, second)
        got = st.weave_transcript_segs(first, second, float(pct))
        exp = S.weave_transcript_segs(first, second, pct)
        if o_c == o_o:
            assert got == exp, (first, second, float(pct))
        n_woven += got != first + " " + second
    assert n_woven > 20  # the alignment branch, not only the concatenation fallback, is exercised


def test_weave_properties(st):
    # nothing in common -> plain concatenation with one space (weaving.rs:217-223)
    assert st.weave_transcript_segs("abc", "xyz", 0.3) == "abc xyz"
    assert st.weave_transcript_segs("", "xyz", 0.3) == " xyz"
    assert st.weave_transcript_segs("abc", "", 0.3) == "abc "
    # min_alignment_score above any reachable score -> concatenation
    assert st.weave_transcript_segs("hello world", "world peace", 0.4, 1e9) == "hello world world peace"
    # multi-byte text survives the byte/char index arithmetic
    a, b = "señor naïve 漢字 fox", "漢字 fox jumps"
    assert st.weave_transcript_segs(a, b, 0.4) == S.weave_transcript_segs(a, b, np.float32(0.4))


def test_is_overlap_silence_and_mean_amplitude_match_oracle(st):
    rng = np.random.default_rng(3)
    for n in (0, 1, 37, 799, 800, 801, 2400):
        for scale in (1e-4, 0.02, 0.3):
            a = (scale * rng.standard_normal(n)).astype(np.float32)
            m = st.mean_amplitude(a)
            assert m == float(S.mean_amplitude(a))
            for mean_amp in (0.0, 0.5 * scale, 2.0 * scale, 8.0 * scale):
                assert st.is_overlap_silence(a, mean_amp) == S.is_overlap_silence(a, np.float32(mean_amp)), (n, scale, mean_amp)
    assert st.is_overlap_silence(np.zeros(0, np.float32), 0.0) is True       # empty overlap counts as silence (weaving.rs:286-288)
    assert st.is_overlap_silence(np.zeros(100, np.float32), 0.0) is False    # strict `<` against 0


def test_window_sequence_matches_oracle_and_tiles_the_buffer(st):
    cases = [(160000, 56000, 16000, 8000), (56000, 56000, 16000, 8000), (56001, 56000, 16000, 8000), (30000, 56000, 16000, 8000),
             (1, 56000, 16000, 8000), (0, 56000, 16000, 8000), (100000, 48000, 0, 0), (123457, 40000, 8000, 4000), (90000, 56000, 16000, 8000)]
    for total, win, lead, trail in cases:
        got = st.window_sequence(total, win, lead, trail)
        exp = S.window_sequence(total, win, lead, trail)
        assert [(g[0], g[1]) for g in got] == [(e[0], e[1]) for e in exp], (total, win)
        assert [float(g[2]) for g in got] == [float(e[2]) for e in exp]
        if total:
            assert got[0][0][0] == 0 and got[-1][0][1] == total
            for (s0, e0), (t0, t1), ov in got:
                assert e0 - s0 <= win and s0 <= t0 <= t1 == e0
            for a, b in zip(got, got[1:]):  # consecutive windows overlap by leading + trailing samples
                assert a[0][1] - b[0][0] == lead + trail
    # the reference's stream parameters: 10 s buffer, 3.5 s windows -> 2 s hop
    w = st.window_sequence(160000, 56000, 16000, 8000)
    assert [x[0][0] for x in w] == [0, 32000, 64000, 96000, 128000]
    with pytest.raises(Exception):
        st.window_sequence(1000, 100, 60, 40)  # window <= contexts would never advance


def test_oracle_buffer_keeps_leading_context_on_overflow():
    """OverlappingAudioBuffer.add_samples (audio.rs:196-242): on overflow the last `leading_context` samples are kept."""
    b = S.OverlappingAudioBuffer(160000, 2.0, 1.0, 0.5)
    x = np.arange(150000, dtype=np.float32)
    b.add_samples(x)
    assert b.length == 150000
    y = np.arange(150000, 170000, dtype=np.float32)
    b.add_samples(y)
    assert b.length == 16000 + 20000
    assert np.array_equal(b.get_window()[:16000], x[-16000:]) and np.array_equal(b.get_window()[16000:], y)
    b2 = S.OverlappingAudioBuffer(1000, 2.0, 1.0, 0.5)
    b2.add_samples(np.ones(1500, np.float32))  # longer than the capacity: truncated, amplitude untouched (audio.rs:236-241)
    assert b2.length == 1000 and b2.mean_amplitude == 0.0


def test_golden_streaming_fixture(st):
    """tests/golden/streaming_golden.json (made by make_streaming_golden.py from the oracle): the C++ behind the C ABI
    reproduces every committed vector."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "streaming_golden.json"), encoding="utf-8"))
    for c in g["weave"]:
        o, s = st.best_alignment(c["first"], c["second"], c["pct"])
        assert abs(s - c["score"]) <= 1e-5 * max(1.0, abs(c["score"]))
        if o == c["overlap"]:
            assert st.weave_transcript_segs(c["first"], c["second"], c["pct"]) == c["woven"], c
    for c in g["windows"]:
        w = st.window_sequence(*c["args"])
        assert [[a[0], a[1], b[0], b[1]] for a, b, _ in w] == c["slices"]
        assert [float(o) for _, _, o in w] == c["overlap"]
    for c in g["silence"]:
        a = np.array(c["audio"], np.float32)
        assert st.mean_amplitude(a) == c["mean_abs"]
        assert st.is_overlap_silence(a, c["mean_amplitude"]) == c["silent"]
