"""GPU tests of the streaming orchestrator: the C++ IncrementalAsr / stream group (csrc/host_stream.cpp)
against the Python oracle's IncrementalAsr (oracle/streaming.py, literal restatement of src/asr/incremental.rs:35-298) driving
the SAME GPU pipeline one request at a time — so what is compared is the orchestration: buffering, window generation, state
carry, silence detection, transcript weaving and token accumulation.  The encoder model is out of scope and injected."""
import numpy as np
import pytest

import streaming as S  # oracle/streaming.py
from conftest import synth_pcm
from test_gpu_pipeline import VOCAB, _write_vocab, stub_encoder

pytestmark = pytest.mark.gpu


def loud_encoder(feats: np.ndarray) -> np.ndarray:
    """The stub encoder with 3x the amplitude: with the synthetic weights this gives a few tokens per second (the plain
    stub gives none on these signals), so the weaving / accumulation branches see real text."""
    return np.ascontiguousarray(3.0 * stub_encoder(feats), dtype=np.float32)


@pytest.fixture(scope="module")
def pipe(amira):
    _write_vocab()
    ctx = amira.Context(device_id=0)
    ctx.load_weights(amira.synthetic_weights(3456))
    p = amira.B200AsrPipeline(ctx, VOCAB, loud_encoder)
    yield p
    p.close()
    ctx.close()


class _Adapter:
    """What the oracle's IncrementalAsr needs from `Arc<dyn AsrPipeline>` (src/asr/pipeline.rs:20-67)."""

    def __init__(self, pipe):
        self.pipe = pipe

    def process_stream_samples(self, samples, state):
        tr = self.pipe.process_stream_samples(samples, state)
        return tr.text, tr.tokens

    def process_batch(self, audio_bytes):
        tr = self.pipe.process_batch(audio_bytes)
        return tr.text, tr.tokens, tr.audio_length_samples, tr.features_length, tr.encoded_length


def _chunks(pcm: np.ndarray, chunk_samples: int):
    return [pcm[i:i + chunk_samples].tobytes() for i in range(0, pcm.size, chunk_samples)]


def test_incremental_asr_equals_the_oracle_chunk_by_chunk(pipe, amira):
    pcm = np.concatenate([synth_pcm(3.5, 900 + i) for i in (1, 3)])
    pcm[16000 * 3:16000 * 4] //= 200  # a quiet second, so that the silence branch of accumulate_transcription is reachable
    ora = S.IncrementalAsr(_Adapter(pipe), lambda: amira.DecoderState.new(1))
    inc = amira.IncrementalAsr(pipe)
    n_nonempty = 0
    for k, ch in enumerate(_chunks(pcm, 16000)):  # 1 s chunks, 7 of them; the buffer passes one 3.5 s window after 4
        exp = ora.process_chunk(ch)
        got = inc.process_chunk(ch)
        assert got == exp, k
        assert inc.tokens == ora.token_ids, k
        assert abs(inc.audio_length() - float(ora.audio_length())) < 1e-6
        n_nonempty += bool(got)
    assert n_nonempty > 0 and ora.n_pipeline_calls > 7, (n_nonempty, ora.n_pipeline_calls)  # later chunks re-decode every window
    inc.clear()
    assert inc.process_chunk(b"") == "" and inc.tokens == []
    inc.close()


def test_buffer_overflow_keeps_context_like_the_oracle(pipe, amira):
    """More audio than the buffer holds (capacity 4 s here): the buffer restarts from the leading context (audio.rs:203-214)."""
    pcm = synth_pcm(6.5, 902)
    ora = S.IncrementalAsr(_Adapter(pipe), lambda: amira.DecoderState.new(1), 1.0, 0.5, 0.25, 4.0)
    inc = amira.IncrementalAsr(pipe, 1.0, 0.5, 0.25, 4.0)
    for k, ch in enumerate(_chunks(pcm, 12000)):
        assert inc.process_chunk(ch) == ora.process_chunk(ch), k
        assert inc.tokens == ora.token_ids
        assert abs(inc.audio_length() - float(ora.audio_length())) < 1e-6
    inc.close()


def test_process_batch_short_and_long(pipe, amira):
    inc = amira.IncrementalAsr(pipe)
    ora = S.IncrementalAsr(_Adapter(pipe), lambda: amira.DecoderState.new(1))
    short = synth_pcm(1.5, 903).tobytes()   # <= chunk_size: straight to AsrPipeline::process_batch (incremental.rs:276-278)
    tr = inc.process_batch(short)
    ref = pipe.process_batch(short)
    assert (tr.text, tr.tokens, tr.features_length, tr.encoded_length) == (ref.text, ref.tokens, ref.features_length, ref.encoded_length)
    long_ = synth_pcm(5.0, 904).tobytes()
    tr = inc.process_batch(long_)
    text, tokens, n, fl, el = ora.process_batch(long_)
    assert (tr.text, tr.tokens, tr.audio_length_samples, tr.features_length, tr.encoded_length) == (text, tokens, n, 0, 0)
    inc.close()


def test_stream_group_equals_one_stream_at_a_time(pipe, amira):
    """64 streams advanced together — one front-end launch + one decode launch per window round — give, per stream, exactly
    what a single IncrementalAsr gives; streams have different chunk sizes, join late and leave early."""
    n = 64
    rng = np.random.default_rng(77)
    pcms = [synth_pcm(float(rng.uniform(1.0, 6.0)), 1000 + i) for i in range(n)]
    sizes = [int(rng.choice([2560, 4000, 8000, 16000])) for _ in range(n)]
    chunks = [_chunks(p, s) for p, s in zip(pcms, sizes)]
    starts = [int(rng.integers(0, 3)) for _ in range(n)]
    grp = amira.StreamGroup(pipe, n)
    ticks = max(st + len(c) for st, c in zip(starts, chunks))
    for t in range(ticks):
        ids = [i for i in range(n) if starts[i] <= t < starts[i] + len(chunks[i])]
        if ids:
            grp.process_chunks(ids, [chunks[i][t - starts[i]] for i in ids])
    calls, rounds = grp.stats()
    assert rounds < calls / 8  # batched: many process_stream_samples calls per launch round
    for i in range(0, n, 7):  # one stream at a time through the same object type
        one = amira.IncrementalAsr(pipe)
        for ch in chunks[i]:
            text = one.process_chunk(ch)
        assert grp.transcript(i) == text, i
        assert grp.tokens(i) == one.tokens, i
        one.close()
    with pytest.raises(amira.AmiraError):
        grp.process_chunks([1, 1], [b"\0\0", b"\0\0"])  # a stream may appear once per call
    grp.close()
