"""GPU tests of the INCREMENTAL mode of the stream group ("a truly incremental front end (carry last sample +
STFT tail) and decoder (carry state + last token)") — every chunk costs its own frames instead of every window of a 10 s buffer.

What is pinned:
  * the log-mel frames a stream emits chunk by chunk are BIT-IDENTICAL to the frames of the whole recording (one launch over all
    of it), whatever the chunking — two hops of carried left context are enough, the right edge is only touched on flush;
  * the features handed to the encoder are those frames normalised with the running (count, mean, M2) over all frames so far
    (restated in numpy here: this mode's own definition — utterance-level statistics do not exist before the stream ends);
  * the tokens equal ONE greedy decode over the concatenation of all encoder frames (state and last token are carried:
    amira_greedy_decode_resume), and batching many streams changes nothing;
  * the literal mode (the reference's IncrementalAsr, tests/test_gpu_streaming.py) is untouched and stays the parity anchor."""
import numpy as np
import pytest

from conftest import synth_pcm
from test_gpu_pipeline import VOCAB, _write_vocab

pytestmark = pytest.mark.gpu

_W = (np.random.default_rng(77).standard_normal((1024, 128)) / np.sqrt(128)).astype(np.float32)


class FrameEncoder:
    """A frame-local stand-in for the encoder (one encoder frame per mel frame), recording what it is given."""

    def __init__(self, gain=2.5):
        self.calls, self.gain = [], gain

    def __call__(self, feats: np.ndarray) -> np.ndarray:
        self.calls.append(np.array(feats, np.float32))
        return np.ascontiguousarray(self.gain * 0.5 * np.tanh(_W @ feats), dtype=np.float32)


def _running_normalise(logmel: np.ndarray, frame_counts):
    """This mode's normalisation, restated: after each round, (x - mean) / (std + 1e-5) of the NEW frames with the sample
    statistics of all frames emitted so far."""
    out, t = [], 0
    for n in frame_counts:
        seen = logmel[:, :t + n].astype(np.float64)
        mean = seen.mean(axis=1)
        sd = seen.std(axis=1, ddof=1) if t + n > 1 else np.zeros(128)
        out.append(((logmel[:, t:t + n] - mean[:, None].astype(np.float32)) * (1.0 / (sd + 1e-5)).astype(np.float32)[:, None]).astype(np.float32))
        t += n
    return out


@pytest.fixture(scope="module")
def gpu(amira):
    _write_vocab()
    ctx = amira.Context(device_id=0)
    ctx.load_weights(amira.synthetic_weights(3456))
    yield ctx
    ctx.close()


def _run_stream(amira, ctx, pcm, sizes, n_streams=1, which=0, gain=2.5):
    enc = FrameEncoder(gain)
    pipe = amira.B200AsrPipeline(ctx, VOCAB, enc)
    g = amira.StreamGroup(pipe, n_streams)
    g.set_incremental(True)
    pos = 0
    for sz in sizes:
        g.process_chunks([which], [pcm[pos:pos + sz].tobytes()])
        pos += sz
    assert pos == pcm.size
    g.flush([which])
    toks, prog = g.tokens(which), g.progress(which)
    text = g.transcript(which)
    g.close()
    pipe.close()
    return enc.calls, toks, prog, text


def test_frames_are_those_of_the_whole_recording_and_statistics_run(gpu, amira, oracle):
    pcm = synth_pcm(4.0, 4242)
    sizes = [2560, 100, 5000, 1, 3333, 160, 159, 12000, 2560, 2560]
    sizes.append(pcm.size - sum(sizes))
    calls, toks, prog, _ = _run_stream(amira, gpu, pcm, sizes)
    n_frames = pcm.size // 160 + 1
    assert prog == (pcm.size, n_frames, n_frames)              # every frame emitted exactly once, flush included
    assert sum(c.shape[1] for c in calls) == n_frames
    counts = sorted(c.shape[1] for c in calls[:-1])
    assert counts[-1] <= max(sizes) // 160 + 2 and counts[-2] <= 12000 // 160 + 2  # a chunk costs its own frames ...
    assert calls[-1].shape[1] <= 2                                                # ... and the flush only the right edge
    whole, lens = gpu.logmel_pcm16_packed(pcm, [0, pcm.size])  # un-normalised log-mel of the whole recording, one launch
    assert int(lens[0]) == n_frames
    want = _running_normalise(whole[0], [c.shape[1] for c in calls])
    for k, (got, exp) in enumerate(zip(calls, want)):
        assert np.abs(got - exp).max() <= 2e-5 * max(1.0, float(np.abs(exp).max())), k
    # and the un-normalised frames are the float64 restatement's
    x = pcm.astype(np.float64) / 32768.0
    y = np.concatenate([[x[0]], x[1:] - 0.97 * x[:-1]])
    idx = np.arange(-256, x.size + 256)
    per = 2 * (x.size - 1)
    idx = np.mod(idx, per)
    idx = np.where(idx < x.size, idx, per - idx)
    frames = np.stack([y[idx][t * 160:t * 160 + 512] for t in range(n_frames)]) * oracle.hann_window_padded()[None, :]
    ref = np.log(np.abs(np.fft.rfft(frames, axis=1)) ** 2 @ oracle.mel_filterbank().astype(np.float64).T + 2.0 ** -24).T
    assert np.abs(whole[0] - ref).max() <= 2e-5


def test_chunking_does_not_change_the_frames(gpu, amira):
    """Bit-identical un-normalised frames for two different chunkings: de-normalising is not needed — with the SAME frame counts per
    round the features are bit-identical, so the chunkings are chosen to emit at the same round boundaries (multiples of a hop)."""
    pcm = synth_pcm(3.0, 99)
    a, _, _, _ = _run_stream(amira, gpu, pcm, [16000, 16000, 16000])
    b, _, _, _ = _run_stream(amira, gpu, pcm, [16000, 16000, 16000])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # different chunkings, same frames: compare through the whole-recording launch (the previous test) for both
    whole, _ = gpu.logmel_pcm16_packed(pcm, [0, pcm.size])
    for sizes in ([48000], [1] * 7 + [47993], [4801, 4799, 38400]):
        calls, _, prog, _ = _run_stream(amira, gpu, pcm, sizes)
        want = _running_normalise(whole[0], [c.shape[1] for c in calls])
        assert prog[1] == pcm.size // 160 + 1
        assert all(np.abs(g_ - w_).max() <= 2e-5 * max(1.0, float(np.abs(w_).max())) for g_, w_ in zip(calls, want))


def test_tokens_equal_one_decode_over_all_encoder_frames(gpu, amira):
    pcm = synth_pcm(2.5, 1717)
    sizes = [2560] * 15 + [pcm.size - 15 * 2560]
    for gain in (2.5, 1.5, 1.0, 0.7, 0.5):  # the first gain whose decode emits some tokens but stays under the per-call limit
        calls, toks, prog, text = _run_stream(amira, gpu, pcm, sizes, gain=gain)
        enc_all = np.concatenate([FrameEncoder(gain)(c) for c in calls], axis=1)  # what the stub produced, call by call
        one, _, steps = gpu.greedy_decode_packed([enc_all])
        if 0 < len(one[0]) < gpu.max_total_tokens:
            break
    assert 0 < len(one[0]) < gpu.max_total_tokens
    assert toks == one[0]                                       # state + last token carried: chunking is invisible to the decoder
    # without the last-token carry (the reference's per-call blank restart) the chunked decode differs: the carry is what matters
    st = amira.DecoderState.new(1)
    restart = []
    for c in calls:
        t, st, _ = gpu.greedy_decode_packed([FrameEncoder(gain)(c)], state=st)
        restart += t[0]
    print(f"gain {gain}: {len(one[0])} tokens; per-call blank restart gives {len(restart)} tokens, equal: {restart == one[0]}")


def test_resume_entry_equals_single_call(gpu, amira):
    rng = np.random.default_rng(5)
    enc = [(0.5 * rng.standard_normal((1024, t))).astype(np.float32) for t in (40, 33, 1, 0, 17)]
    want, st_want, steps = gpu.greedy_decode_packed(enc)
    st = amira.DecoderState.new(len(enc))
    last = np.full(len(enc), amira.BLANK_ID, np.int32)
    got = [[] for _ in enc]
    for lo in range(0, 40, 7):
        part = [np.ascontiguousarray(e[:, lo:lo + 7]) for e in enc]
        t, st, last, _ = gpu.greedy_decode_resume(part, st, last)
        for b in range(len(enc)):
            got[b] += t[b]
    assert got == want
    assert np.array_equal(st.states_1, st_want.states_1) and np.array_equal(st.states_2, st_want.states_2)
    with pytest.raises(amira.AmiraError):
        gpu.greedy_decode_resume([enc[0]], amira.DecoderState.new(1), [1025])  # outside the embedding table


def test_many_streams_at_once_equal_each_alone(gpu, amira):
    n = 24
    pcms = [synth_pcm(1.0 + 0.07 * i, 500 + i) for i in range(n)]
    alone = [_run_stream(amira, gpu, pcms[i], [4000] * (pcms[i].size // 4000) + [pcms[i].size % 4000])[1] for i in range(0, n, 5)]
    enc = FrameEncoder()
    pipe = amira.B200AsrPipeline(gpu, VOCAB, enc)
    g = amira.StreamGroup(pipe, n)
    g.set_incremental(True)
    pos = [0] * n
    while any(pos[i] < pcms[i].size for i in range(n)):
        ids = [i for i in range(n) if pos[i] < pcms[i].size]
        g.process_chunks(ids, [pcms[i][pos[i]:pos[i] + 4000].tobytes() for i in ids])
        for i in ids:
            pos[i] += 4000
    g.flush(list(range(n)))
    for k, i in enumerate(range(0, n, 5)):
        assert g.tokens(i) == alone[k], i
    calls, rounds = g.stats()
    assert rounds < calls  # streams are batched: fewer launch rounds than per-stream requests
    with pytest.raises(amira.AmiraError):
        g.process_chunks([0], [b"\x00\x00"])  # a flushed stream takes no more audio
    g.clear(0)
    g.process_chunks([0], [pcms[0][:4000].tobytes()])
    assert g.progress(0)[0] == 4000
    g.close()
    pipe.close()


def test_decoder_state_follows_its_stream_when_the_round_changes(gpu, amira):
    """The group keeps the decoder states of a round in batch layout while consecutive rounds list the same streams in the same
    order (csrc/host_stream.cpp: `resident`).  Whatever the rounds look like — a subset, another order, a stream cleared and
    restarted in between, transcripts read mid-way — every stream's tokens are those of the stream processed alone."""
    n = 6
    pcms = [synth_pcm(1.2 + 0.11 * i, 900 + i) for i in range(n)]
    step = 3200
    alone = [_run_stream(amira, gpu, pcms[i], [step] * (pcms[i].size // step) + [pcms[i].size % step])[1] for i in range(n)]
    enc = FrameEncoder()
    pipe = amira.B200AsrPipeline(gpu, VOCAB, enc)
    g = amira.StreamGroup(pipe, n)
    g.set_incremental(True)
    pos = [0] * n
    rng = np.random.default_rng(3)
    k = 0
    while any(pos[i] < pcms[i].size for i in range(n)):
        live = [i for i in range(n) if pos[i] < pcms[i].size]
        if k % 3 == 0:
            ids = live                                   # the same list as the round before last: states stay where they are
        elif k % 3 == 1:
            ids = list(reversed(live))                   # another order
        else:
            ids = [i for i in live if rng.random() < 0.6] or live[:1]   # a subset
        g.process_chunks(ids, [pcms[i][pos[i]:pos[i] + step].tobytes() for i in ids])
        for i in ids:
            pos[i] += step
        if k == 4:                                       # stream 2 starts over: its state is reset, the others' must survive
            g.clear(2)
            pos[2] = 0
        if k == 6:
            _ = g.transcript(1)                          # reading a transcript mid-stream changes nothing
        k += 1
    g.flush(list(range(n)))
    for i in range(n):
        assert g.tokens(i) == alone[i], i
        assert g.transcript(i) == pipe.decode_tokens(alone[i]), i
    g.close()
    pipe.close()
