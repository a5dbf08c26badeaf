"""GPU tests of the host pipeline above the C ABI: the C++ mirror of the reference's `AsrPipeline` trait
(src/asr/pipeline.rs:20-67, process_audio_zero_copy :269-380) and the request micro-batcher.  The encoder model is out
of scope and injected; here it is a fixed numpy function so that every path sees the same encoder."""
import os
import threading

import numpy as np
import pytest

from conftest import synth_pcm

pytestmark = pytest.mark.gpu
NEAR_TIE = 2e-4
VOCAB = "/tmp/amira_test_vocab.txt"


def _write_vocab():
    with open(VOCAB, "w", encoding="utf-8") as f:
        for i in range(1024):
            f.write(("▁w%d %d\n" if i % 3 == 0 else "p%d %d\n") % (i, i))
        f.write("<blk> 1024\n")


_rng = np.random.default_rng(123)
_W = (_rng.standard_normal((1024, 128)) / np.sqrt(128)).astype(np.float32)


def stub_encoder(feats: np.ndarray) -> np.ndarray:
    """[128, L] -> [1024, T], T = the 8x subsampled length of the real encoder (L <- (L-1)//2+1 three times)."""
    L = feats.shape[1]
    T = L
    for _ in range(3):
        T = (T - 1) // 2 + 1 if T > 0 else 0
    if T == 0:
        return np.zeros((1024, 0), np.float32)
    pooled = np.stack([feats[:, 8 * t:8 * t + 8].mean(axis=1) for t in range(T)], axis=1)
    return np.ascontiguousarray(0.5 * np.tanh(_W @ pooled), dtype=np.float32)


@pytest.fixture(scope="module")
def pipe(amira):
    _write_vocab()
    ctx = amira.Context(device_id=0)
    ctx.load_weights(amira.synthetic_weights(3456))
    p = amira.B200AsrPipeline(ctx, VOCAB, stub_encoder)
    yield p, ctx
    p.close()
    ctx.close()


def test_process_batch_equals_the_composed_primitives_and_the_oracle(pipe, amira, oracle):
    p, ctx = pipe
    model = oracle.Model(blob=amira.synthetic_weights(3456))
    vocab = oracle.Vocabulary.load_from_file(VOCAB)
    for i, secs in enumerate((2.0, 0.6, 3.3)):
        pcm = synth_pcm(secs, 40 + i)
        tr = p.process_batch(pcm.tobytes())
        # same thing through the library's own primitives
        feats, lens = ctx.preprocess_pcm16(pcm, [0, pcm.size])
        enc = stub_encoder(feats[0, :, :int(lens[0])])
        toks, _, _ = ctx.greedy_decode(enc[None], [enc.shape[1]])
        assert tr.tokens == toks[0]
        assert (tr.audio_length_samples, tr.features_length, tr.encoded_length) == (pcm.size, int(lens[0]), enc.shape[1])
        assert tr.text == vocab.decode_tokens(tr.tokens) == p.decode_tokens(tr.tokens)
        # and through the CPU oracle end to end (features differ by <= 1e-4, so only near-ties may flip a token)
        ref_f, L = oracle.preprocess(pcm.astype(np.float32) / 32768.0, "f64")
        ref_enc = stub_encoder(ref_f[:, :L])
        r = oracle.greedy_decode(ref_enc, ref_enc.shape[1], model)
        assert r.tokens == tr.tokens or r.margins.min() < 50 * NEAR_TIE


def test_stream_chunks_carry_the_lstm_state(pipe, amira):
    p, ctx = pipe
    pcm = synth_pcm(1.92, 77)
    st = amira.DecoderState.new(1)
    toks_stream = []
    for c in range(4):  # four 480 ms chunks; token history is per call, the state is carried (pipeline.rs:384-401)
        chunk = pcm[c * 7680:(c + 1) * 7680]
        tr = p.process_stream_chunk(chunk.tobytes(), st)
        feats, lens = ctx.preprocess_pcm16(chunk, [0, chunk.size])
        enc = stub_encoder(feats[0, :, :int(lens[0])])
        if c == 0:
            ref_state = amira.DecoderState.new(1)
        ref_toks, ref_state, _ = ctx.greedy_decode(enc[None], [enc.shape[1]], state=ref_state)
        assert tr.tokens == ref_toks[0]
        toks_stream += tr.tokens
    assert np.array_equal(st.states_1, ref_state.states_1) and np.array_equal(st.states_2, ref_state.states_2)
    assert np.abs(st.states_1).max() > 0


def test_odd_length_and_empty_requests(pipe):
    p, _ = pipe
    pcm = synth_pcm(0.5, 9)
    odd = pcm.tobytes() + b"\x05"  # trailing byte: bytes_to_f32_optimized rule (performance_opts.rs:26-30)
    tr = p.process_batch(odd)
    assert tr.audio_length_samples == pcm.size + 1
    tr0 = p.process_batch(b"")
    assert tr0.tokens == [] and tr0.text == ""


def test_micro_batcher_equals_one_by_one(pipe, amira):
    """Concurrent process_batch calls are coalesced into few launches; every caller gets exactly the one-by-one result."""
    p, ctx = pipe
    utts = [synth_pcm(float(0.4 + 0.13 * i), 500 + i).tobytes() for i in range(24)]
    want = [p.process_batch(u) for u in utts]
    batcher = amira.Batcher(p, max_batch=16, max_wait_us=20000)
    got = [None] * len(utts)
    errs = []

    def work(i):
        try:
            got[i] = batcher.process_batch(utts[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(utts))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs
    for g, w in zip(got, want):
        assert g.tokens == w.tokens and g.text == w.text
        assert (g.audio_length_samples, g.features_length, g.encoded_length) == (w.audio_length_samples, w.features_length, w.encoded_length)
    n_req, n_batches = batcher.stats()
    assert n_req == len(utts) and n_batches < len(utts)
    # empty request through the batcher takes the single-request path
    assert batcher.process_batch(b"").tokens == []
    batcher.close()


def test_two_gpus_in_one_process(amira):
    """One process, one context per GPU: per-device kernel attributes, tables and weights — both devices
    must give the results of device 0.  Skipped on single-GPU boxes."""
    if amira.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(5)
    pcms = [synth_pcm(float(rng.uniform(0.5, 2.0)), 700 + i) for i in range(40)]
    offs = np.zeros(len(pcms) + 1, np.int64)
    offs[1:] = np.cumsum([p.size for p in pcms])
    pcm = np.concatenate(pcms)
    enc = (0.5 * rng.standard_normal((40, 1024, 20))).astype(np.float32)
    outs = []
    for dev in (1, 0):  # device 1 first: nothing may depend on device 0 having been initialised
        with amira.Context(device_id=dev) as c:
            c.load_weights(amira.synthetic_weights(3456))
            feats, lens = c.preprocess_pcm16(pcm, offs)
            toks, st, steps = c.greedy_decode(enc)
            outs.append((feats, lens, toks, st.states_1, steps))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2] and np.array_equal(outs[0][3], outs[1][3]) and np.array_equal(outs[0][4], outs[1][4])


def test_forked_lanes_share_weights_and_run_concurrently(amira, oracle):
    """amira_ctx_fork: lanes of one context decode concurrently from several threads with ONE copy of the weights; a reload
    through any lane is seen by all of them; lanes can be destroyed in any order (the parent first)."""
    import threading
    rng = np.random.default_rng(41)
    enc = (0.5 * rng.standard_normal((6, 1024, 20))).astype(np.float32)
    root = amira.Context(device_id=0)
    root.load_weights(amira.synthetic_weights(3456))
    want, _, want_steps = root.greedy_decode(enc)
    lanes = [root.fork() for _ in range(3)]
    got = [None] * 3

    def run(i):
        for _ in range(4):
            got[i] = lanes[i].greedy_decode(enc)

    ths = [threading.Thread(target=run, args=(i,)) for i in range(3)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for i in range(3):
        assert got[i][0] == want and got[i][2].tolist() == want_steps.tolist()
    # front end on a lane
    pcm = (rng.standard_normal(16000) * 3000).astype(np.int16)
    f0, _ = root.preprocess_pcm16(pcm, [0, pcm.size])
    f1, _ = lanes[1].preprocess_pcm16(pcm, [0, pcm.size])
    assert np.array_equal(f0, f1)
    # a reload through one lane reaches the others; the parent goes first
    lanes[2].load_weights(amira.synthetic_weights(777))
    other, _, _ = lanes[0].greedy_decode(enc)
    root.close()
    again, _, _ = lanes[1].greedy_decode(enc)
    assert other == again
    model = oracle.Model(blob=amira.synthetic_weights(777))
    r = oracle.greedy_decode(enc[0], 20, model)
    assert again[0] == r.tokens or r.margins.min() < 2e-4
    with pytest.raises(amira.AmiraError):
        lanes[0].stream_open()  # stream slots stay with the parent
    for lane in lanes:
        lane.close()


_IPC_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import amira_b200 as A
handle = bytes.fromhex(sys.argv[2]); B, T = int(sys.argv[3]), int(sys.argv[4])
with A.Context(device_id=0) as ctx:
    ctx.load_weights(A.synthetic_weights(3456))
    enc = ctx.ipc_import(handle)                      # the region the parent process exported
    tok = np.zeros((B, 200), np.int32); nt = np.zeros(B, np.int32)
    ctx.greedy_decode_raw(enc, B, T, None, tok.ctypes.data, nt.ctypes.data)
    ctx.ipc_close(enc)
    print("TOKENS", ";".join(",".join(str(x) for x in tok[b, :nt[b]]) for b in range(B)))
"""


def test_device_handoff_across_processes(amira, tmp_path):
    """src/cuda/cuda_helper.cu:63-183 precedent: a device region allocated and exported by one process (amira_device_alloc +
    amira_ipc_export) is opened by another (amira_ipc_import) and decoded in place — encoder outputs never touch host memory."""
    import subprocess
    import sys
    from cuda import cudart
    rng = np.random.default_rng(51)
    B, T = 5, 16
    enc = (0.5 * rng.standard_normal((B, 1024, T))).astype(np.float32)
    with amira.Context(device_id=0) as ctx:
        ctx.load_weights(amira.synthetic_weights(3456))
        want, _, _ = ctx.greedy_decode(enc)
        region = ctx.device_alloc(enc.nbytes)
        (err,) = cudart.cudaMemcpy(region, enc.ctypes.data, enc.nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice)
        assert int(err) == 0
        handle = ctx.ipc_export(region)
        assert len(handle) == 64 and any(handle)
        # in-process: the exported region is used in place through its device pointer
        tok = np.zeros((B, 200), np.int32)
        nt = np.zeros(B, np.int32)
        ctx.greedy_decode_raw(region, B, T, None, tok.ctypes.data, nt.ctypes.data)
        assert [tok[b, :nt[b]].tolist() for b in range(B)] == want
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        script = tmp_path / "child.py"
        script.write_text(_IPC_CHILD)
        r = subprocess.run([sys.executable, str(script), root, handle.hex(), str(B), str(T)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("TOKENS")][0][7:]
        got = [[int(x) for x in part.split(",")] if part else [] for part in line.split(";")]
        assert got == want
        ctx.device_free(region)
