"""GPU parity of the persistent greedy-decode kernel and the decoder_joint contract op against the CPU oracle
(literal restatement of src/asr/decoder_optimized.rs:54-200 around an fp32 LSTM/joint step).

Token IDs must be bit-exact except documented argmax near-ties: a stream may diverge from the oracle only at a step
whose oracle top-1/top-2 logit margin is below NEAR_TIE (fp32 re-association noise is ~1e-5 on logits of O(1))."""
import numpy as np
import pytest

from conftest import calibrated_weights

pytestmark = pytest.mark.gpu
NEAR_TIE = 2e-4


ENGINES = {1: "fp32 persistent kernel (numerics anchor)", 4: "tcgen05 weight-stationary dataflow (default)"}


@pytest.fixture(scope="module", params=[1, 4], ids=["fp32", "ws"])
def ctx(request, amira):
    """One context per decode engine: every test in this module runs against both."""
    c = amira.Context(device_id=0, decode_engine=request.param)
    c.engine = request.param
    yield c
    c.close()


@pytest.fixture(scope="module")
def model(oracle, ctx):
    blob = calibrated_weights(oracle)
    ctx.load_weights(blob)
    return oracle.Model(blob=blob)


def _check_tokens(oracle, model, enc, lens, got_tokens, got_steps, states=None):
    n_div = 0
    for b in range(enc.shape[0]):
        L = int(lens[b])
        st = None if states is None else (states[0][:, b], states[1][:, b])
        r = oracle.greedy_decode(np.ascontiguousarray(enc[b, :, :L]), L, model, states=st)
        assert r.rc == 0
        if got_tokens[b] == r.tokens:
            assert got_steps[b] == r.n_steps
            continue
        # a diverging stream must contain a near-tie step in the oracle (documented exception)
        n_div += 1
        assert r.margins.min() < NEAR_TIE, (b, float(r.margins.min()))
    return n_div


def test_decoder_joint_contract_op(ctx, oracle, model):
    rng = np.random.default_rng(5)
    B, T, U = 3, 4, 3
    enc = (0.5 * rng.standard_normal((B, 1024, T))).astype(np.float32)
    tg = np.array([[1024, 5, 17], [1024, 1000, 3], [7, 7, 7]], np.int32)
    s1 = (0.1 * rng.standard_normal((2, B, 640))).astype(np.float32)
    s2 = (0.1 * rng.standard_normal((2, B, 640))).astype(np.float32)
    import amira_b200 as A
    out, pl, st = ctx.decoder_joint(enc, tg, state=A.DecoderState(s1, s2))
    assert out.shape == (B, U, T, 1030) and pl.tolist() == [U] * B
    for b in range(B):
        ref, r1, r2 = model.decoder_joint(enc[b], tg[b], s1[:, b:b + 1], s2[:, b:b + 1])
        assert np.abs(out[b] - ref).max() < 1e-4
        assert np.abs(st.states_1[:, b] - r1.reshape(2, 640)).max() < 5e-5
        assert np.abs(st.states_2[:, b] - r2.reshape(2, 640)).max() < 5e-5
        assert np.array_equal(out[b].reshape(U, T, 1030).argmax(-1), ref.argmax(-1))


def test_decoder_joint_rejects_out_of_table_target(ctx, amira, model):
    enc = np.zeros((1, 1024, 1), np.float32)
    with pytest.raises(amira.AmiraError) as e:
        ctx.decoder_joint(enc, np.array([[1024, 1027]], np.int32))
    assert e.value.code == 6  # "Decode step failed"


def test_greedy_batch_matches_oracle(ctx, oracle, model):
    rng = np.random.default_rng(2345)
    B, T = 24, 40
    enc = (0.5 * rng.standard_normal((B, 1024, T))).astype(np.float32)
    lens = rng.integers(1, T + 1, size=B)
    lens[0], lens[1], lens[2] = T, 0, 1
    toks, st, steps = ctx.greedy_decode(enc, lens)
    assert toks[1] == [] and steps[1] == 0
    n_div = _check_tokens(oracle, model, enc, lens, toks, steps)
    assert n_div <= 1
    # final state of an un-diverged stream equals the oracle's (carried unconditionally, also on blank)
    r = oracle.greedy_decode(np.ascontiguousarray(enc[0, :, :T]), T, model)
    if r.tokens == toks[0]:
        assert np.abs(st.states_1[:, 0] - r.states_1.reshape(2, 640)).max() < 1e-4
        assert np.abs(st.states_2[:, 0] - r.states_2.reshape(2, 640)).max() < 1e-4
    emitted = sum(len(t) for t in toks)
    assert 0 < emitted < int(lens.sum()) * 30


def test_greedy_carried_state_streaming(ctx, oracle, model, amira):
    """Chunked decode with carried LSTM state == the oracle run chunk by chunk (tokens history is per call)."""
    rng = np.random.default_rng(77)
    enc = (0.5 * rng.standard_normal((2, 1024, 12))).astype(np.float32)
    st = amira.DecoderState.new(2)
    o_s = [(np.zeros((2, 1, 640), np.float32), np.zeros((2, 1, 640), np.float32)) for _ in range(2)]
    for c in range(4):
        chunk = np.ascontiguousarray(enc[:, :, 3 * c:3 * c + 3])
        toks, st, _ = ctx.greedy_decode(chunk, [3, 3], state=st)
        for b in range(2):
            r = oracle.greedy_decode(chunk[b], 3, model, states=o_s[b])
            o_s[b] = (r.states_1, r.states_2)
            assert toks[b] == r.tokens or r.margins.min() < NEAR_TIE
    # resident slots give the same answer as caller-owned state
    s0, s1 = ctx.stream_open(), ctx.stream_open()
    got = [[], []]
    for c in range(4):
        chunk = np.ascontiguousarray(enc[:, :, 3 * c:3 * c + 3])
        toks, _ = ctx.stream_decode([s0, s1], chunk)
        got[0] += toks[0]
        got[1] += toks[1]
    fin = ctx.stream_get_state(s1)
    assert np.abs(fin.states_1[:, 0] - st.states_1[:, 1]).max() < 1e-6
    ctx.stream_close(s0)
    ctx.stream_close(s1)


@pytest.mark.parametrize("engine", [1, 4])
def test_limits_max_symbols_and_max_total(oracle, amira, engine):
    """Mock-model KATs of the reference loop (decoder_optimized.rs:331-366 derived): a model that never predicts blank
    emits exactly max_symbols tokens per frame and stops at max_total_tokens."""
    blob = calibrated_weights(oracle, blank_bias=-50.0)  # blank never wins
    model = oracle.Model(blob=blob)
    rng = np.random.default_rng(3)
    enc = (0.5 * rng.standard_normal((2, 1024, 9))).astype(np.float32)
    with amira.Context(device_id=0, decode_engine=engine) as c:
        c.load_weights(blob)
        toks, _, steps = c.greedy_decode(enc, [3, 9])
        assert len(toks[0]) == 90 and steps[0] == 90       # 3 frames x 30 symbols
        assert len(toks[1]) == 200 and steps[1] == 200     # capped inside frame 6
        for b, L in ((0, 3), (1, 9)):
            r = oracle.greedy_decode(np.ascontiguousarray(enc[b, :, :L]), L, model)
            assert toks[b] == r.tokens or r.margins.min() < NEAR_TIE
    with amira.Context(device_id=0, max_symbols_per_step=2, max_total_tokens=5, decode_engine=engine) as c:
        c.load_weights(blob)
        toks, _, steps = c.greedy_decode(enc, [2, 9])
        assert len(toks[0]) == 4 and len(toks[1]) == 5 and steps[1] == 5


@pytest.mark.parametrize("engine", [1, 4])
def test_all_blank_updates_state_every_frame(oracle, amira, engine):
    blob = calibrated_weights(oracle, blank_bias=50.0)  # blank always wins
    model = oracle.Model(blob=blob)
    rng = np.random.default_rng(4)
    enc = (0.5 * rng.standard_normal((1, 1024, 7))).astype(np.float32)
    with amira.Context(device_id=0, decode_engine=engine) as c:
        c.load_weights(blob)
        toks, st, steps = c.greedy_decode(enc)
        assert toks[0] == [] and steps[0] == 7
        r = oracle.greedy_decode(enc[0], 7, model)
        assert r.n_steps == 7 and np.abs(st.states_1[:, 0] - r.states_1.reshape(2, 640)).max() < 1e-5


@pytest.mark.parametrize("engine", [1, 4])
def test_out_of_table_argmax_fails_the_stream(oracle, amira, engine):
    """Flat argmax over all 1030 outputs (zero_copy.rs:190-232) can pick 1025..1029; the reference's next step then
    fails ("Decode step failed").  Same here: n_tokens = -1 for that stream, status AMIRA_ERR_DECODE_STEP."""
    m = oracle.Model(seed=3456)
    blob = m.blob.copy()
    blob[-1030:][1027] += 100.0
    rng = np.random.default_rng(9)
    enc = (0.5 * rng.standard_normal((1, 1024, 3))).astype(np.float32)
    r = oracle.greedy_decode(enc[0], 3, oracle.Model(blob=blob))
    assert r.rc == -1
    with amira.Context(device_id=0, decode_engine=engine) as c:
        c.load_weights(blob)
        toks, _, _ = c.greedy_decode(enc, allow_failed=True)
        assert toks[0] is None
        with pytest.raises(amira.AmiraError) as e:
            c.greedy_decode(enc)
        assert e.value.code == 6


def test_no_weights_is_an_error(amira):
    with amira.Context(device_id=0) as c:
        with pytest.raises(amira.AmiraError) as e:
            c.greedy_decode(np.zeros((1, 1024, 1), np.float32))
        assert e.value.code == 7


def test_golden_fixture(ctx, amira):
    """Committed golden tokens (tests/golden/make_golden.py, produced by the oracle); min oracle margin there is 7e-3,
    far above NEAR_TIE, so the comparison is exact."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "decode_golden.npz"))
    ctx.load_weights(amira.synthetic_weights(int(g["seed"])))
    toks, _, steps = ctx.greedy_decode(g["enc"].astype(np.float32), g["lens"])
    for b in range(len(toks)):
        assert toks[b] == g["tokens"][b, :int(g["n_tokens"][b])].tolist()
        assert steps[b] == int(g["n_steps"][b])
    assert float(g["min_margin"].min()) > 10 * NEAR_TIE


def test_full_size_batch_256_properties(ctx, amira):
    """BASELINE config 3 (256 streams, T = 126, limits 30/200): size-independent properties — batch invariance (every
    stream decodes as it does alone / in a small batch), limits respected, steps = frames visited + tokens."""
    rng = np.random.default_rng(2345)
    B, T = 256, 126
    base = (0.5 * rng.standard_normal((8, 1024, T))).astype(np.float32)
    enc = np.ascontiguousarray(base[np.arange(B) % 8])
    lens = np.full(B, T, np.int64)
    lens[8:16] = 60
    ctx.load_weights(amira.synthetic_weights(3456))
    toks, _, steps = ctx.greedy_decode(enc, lens)
    small, _, ssteps = ctx.greedy_decode(base, np.full(8, T, np.int64))
    for b in range(B):
        assert len(toks[b]) <= 200 and 0 <= steps[b] <= 30 * lens[b]
        if lens[b] == T:
            assert toks[b] == small[b % 8] and steps[b] == ssteps[b % 8]
        if len(toks[b]) < 200:
            assert steps[b] >= lens[b] + len(toks[b]) - 29  # every frame ends with a blank step or the symbol limit


def test_pipelined_host_upload_equals_resident_decode(amira):
    """Encoder outputs in host memory are uploaded chunk by chunk, overlapped with the projection of the chunks already on
    the device (>= 64 streams => several chunks); the tokens must equal the device-resident decode of the same batch."""
    import os
    import torch
    rng = np.random.default_rng(21)
    B, T = 96, 24
    enc = (0.5 * rng.standard_normal((B, 1024, T))).astype(np.float32)
    lens = rng.integers(1, T + 1, size=B).astype(np.int64)
    with amira.Context(device_id=0) as c:
        c.load_weights(amira.synthetic_weights(3456))
        os.environ["AMIRA_FORCE_CHUNKS"] = "3"  # the library chunks by bytes (>= 24 MB per chunk): force the chunked path here
        try:
            toks_h, _, steps_h = c.greedy_decode(enc, lens)
            ragged = [np.ascontiguousarray(enc[b, :, :int(lens[b])]) for b in range(B)]
            toks_k, _, steps_k = c.greedy_decode_packed(ragged)
        finally:
            os.environ.pop("AMIRA_FORCE_CHUNKS", None)
        assert toks_k == toks_h and steps_k.tolist() == steps_h.tolist()
        enc_d = torch.from_numpy(enc).cuda()
        tok_d = torch.zeros((B, c.max_total_tokens), dtype=torch.int32, device="cuda")
        nt_d = torch.zeros(B, dtype=torch.int32, device="cuda")
        ns_d = torch.zeros(B, dtype=torch.int32, device="cuda")
        c.greedy_decode_raw(enc_d.data_ptr(), B, T, lens, tok_d.data_ptr(), nt_d.data_ptr(), ns_d.data_ptr())
        torch.cuda.synchronize()
        nt, tk, ns = nt_d.cpu().numpy(), tok_d.cpu().numpy(), ns_d.cpu().numpy()
        for b in range(B):
            assert toks_h[b] == tk[b, :nt[b]].tolist(), b
            assert steps_h[b] == ns[b]


def test_packed_encoder_outputs_equal_padded(ctx, amira):
    """amira_greedy_decode_packed: ragged per-stream [1024][T_b] blocks (the reference's per-request encoder tensors) give
    the tokens, steps and final states of the padded [B][1024][T] batch — host pipeline (96 streams => several upload
    chunks) and device-resident, including a zero-length stream."""
    import torch
    rng = np.random.default_rng(22)
    B, T = 96, 24
    enc = (0.5 * rng.standard_normal((B, 1024, T))).astype(np.float32)
    lens = rng.integers(1, T + 1, size=B).astype(np.int64)
    lens[5] = 0
    ctx.load_weights(amira.synthetic_weights(3456))
    toks, st, steps = ctx.greedy_decode(enc, lens)
    ragged = [np.ascontiguousarray(enc[b, :, :int(lens[b])]) for b in range(B)]
    if ctx.engine == 1:
        with pytest.raises(amira.AmiraError):
            ctx.greedy_decode_packed(ragged)
        return
    toks_k, st_k, steps_k = ctx.greedy_decode_packed(ragged)
    assert toks_k == toks and steps_k.tolist() == steps.tolist()
    assert np.array_equal(st_k.states_1, st.states_1) and np.array_equal(st_k.states_2, st.states_2)
    # device-resident packed buffer with gaps between blocks
    eoff = np.zeros(B + 1, np.int64)
    eoff[0] = 3
    for b in range(B):
        eoff[b + 1] = eoff[b] + 1024 * int(lens[b]) + (b % 5)
    flat = np.zeros(int(eoff[-1]), np.float32)
    for b in range(B):
        flat[eoff[b]:eoff[b] + 1024 * int(lens[b])] = ragged[b].reshape(-1)
    enc_d = torch.from_numpy(flat).cuda()
    tok_d = torch.zeros((B, ctx.max_total_tokens), dtype=torch.int32, device="cuda")
    nt_d = torch.zeros(B, dtype=torch.int32, device="cuda")
    ns_d = torch.zeros(B, dtype=torch.int32, device="cuda")
    ctx.greedy_decode_packed_raw(enc_d.data_ptr(), eoff, B, lens, tok_d.data_ptr(), nt_d.data_ptr(), ns_d.data_ptr())
    torch.cuda.synchronize()
    nt, tk, ns = nt_d.cpu().numpy(), tok_d.cpu().numpy(), ns_d.cpu().numpy()
    for b in range(B):
        assert toks[b] == tk[b, :nt[b]].tolist(), b
        assert steps[b] == ns[b]


def test_streaming_tick_1024_slots(amira):
    """BASELINE config 4 shape: 1024 resident stream slots, T = 3 encoder frames per tick; two ticks through the slots equal
    one decode with caller-carried state."""
    rng = np.random.default_rng(31)
    n, T = 1024, 3
    enc = (0.5 * rng.standard_normal((2, n, 1024, T))).astype(np.float32)
    with amira.Context(device_id=0) as c:
        c.load_weights(amira.synthetic_weights(3456))
        slots = [c.stream_open() for _ in range(n)]
        assert len(set(slots)) == n
        st = amira.DecoderState.new(n)
        for tick in range(2):
            got, _ = c.stream_decode(slots, enc[tick])
            ref, st, _ = c.greedy_decode(enc[tick], [T] * n, state=st)
            assert got == ref
        with pytest.raises(amira.AmiraError):
            c.stream_decode([slots[0], slots[0]], enc[0, :2])  # duplicate slot in one tick
        for s in slots:
            c.stream_close(s)


def test_superseded_engines_are_rejected(amira):
    """Engines 2 and 3 of round 1 are no longer part of the product library."""
    for eng in (2, 3, 5):
        with pytest.raises(amira.AmiraError) as e:
            amira.Context(device_id=0, decode_engine=eng)
        assert e.value.code == 1


_TRAP_CHILD = r"""
import os, sys, time, numpy as np
sys.path.insert(0, sys.argv[1])
import amira_b200 as A
enc = (0.5 * np.random.default_rng(3).standard_normal((2, 1024, 6))).astype(np.float32)
blob = A.synthetic_weights(3456)
ctx = A.Context(device_id=0); ctx.load_weights(blob)
want, _, _ = ctx.greedy_decode(enc)
print("TOKENS", want)
if sys.argv[2] == "plain":
    sys.exit(0)
os.environ["AMIRA_DEBUG_FORCE_TRAP"] = "1"           # the kernel takes its watchdog exit
try:
    ctx.greedy_decode(enc); print("NO ERROR"); sys.exit(1)
except A.AmiraError as e:
    print("TRAP status", e.code, e.message[:60])
    assert e.code == 3
os.environ.pop("AMIRA_DEBUG_FORCE_TRAP")
try:
    ctx.greedy_decode(enc); print("NOT STICKY"); sys.exit(1)   # the CUDA context of the process is poisoned
except A.AmiraError as e:
    print("sticky status", e.code)
ctx.close()
# in-process recovery: destroy, reset, re-create.  Whether the driver hands the device back to the SAME process right away depends on
# its compute mode (an exclusive-process GPU stays 'busy' until the faulted context is torn down); the portable recovery is a
# process restart, which the parent test checks.
recovered = False
for attempt in range(20):
    try:
        A.device_reset(0)
        ctx = A.Context(device_id=0); ctx.load_weights(blob)
        got, _, _ = ctx.greedy_decode(enc)
        recovered = got == want
        break
    except A.AmiraError as e:
        time.sleep(0.25)
print("IN-PROCESS RECOVERY", recovered)
print("DONE")
"""


def test_watchdog_trap_is_reported_and_contained(tmp_path):
    """A watchdog trap in the persistent kernel surfaces as AMIRA_ERR_UNKNOWN and poisons every context of the device IN THAT
    PROCESS (sticky CUDA error).  Recovery (INTEGRATION.md 'failure and recovery'): amira_ctx_destroy + amira_device_reset +
    amira_ctx_create where the driver allows it, otherwise a process restart — the GPU itself is not wedged: a fresh process
    decodes the same tokens right after.  Runs in child processes: the sticky error must not reach this one."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "trap_child.py"
    script.write_text(_TRAP_CHILD)
    ref = subprocess.run([sys.executable, str(script), root, "plain"], capture_output=True, text=True, timeout=300)
    assert ref.returncode == 0, ref.stderr[-1500:]
    r = subprocess.run([sys.executable, str(script), root, "trap"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "TRAP status 3" in r.stdout and "sticky status" in r.stdout and "DONE" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
    print([l for l in r.stdout.splitlines() if l.startswith("IN-PROCESS")])
    after = subprocess.run([sys.executable, str(script), root, "plain"], capture_output=True, text=True, timeout=300)
    assert after.returncode == 0, after.stderr[-1500:]
    tok = lambda out: [l for l in out.splitlines() if l.startswith("TOKENS")][0]
    assert tok(after.stdout) == tok(ref.stdout) == tok(r.stdout)


def test_literal_refeed_loop_over_the_contract_op(amira, oracle):
    """The reference's loop AS WRITTEN re-sends [blank] ++ tokens of the call on every step and takes the argmax over the whole
    flattened [U,1030] output (src/asr/decoder_optimized.rs:140-163, src/triton/model.rs:713; SURVEY finding 5).  Here the
    oracle's restatement of that loop drives the GPU contract op (amira_decoder_joint, one call per step, [1,1024,1] frame, U
    targets, carried state — exactly the RPC of src/asr/pipeline.rs:323-348) and must produce what it produces around the
    CPU model: same tokens (flat indices included), same step count, same failure behaviour."""
    blob = calibrated_weights(oracle, blank_bias=1.0)  # a blank-shy model: the refeed rows get to compete
    model = oracle.Model(blob=blob)
    rng = np.random.default_rng(21)
    with amira.Context(device_id=0) as c:
        c.load_weights(blob)
        calls = []

        def gpu_step(frame, targets, s1, s2):
            calls.append(len(targets))
            try:
                out, _, st = c.decoder_joint(frame.reshape(1, 1024, 1), targets.reshape(1, -1),
                                             state=amira.DecoderState(s1.reshape(2, 1, 640), s2.reshape(2, 1, 640)))
            except amira.AmiraError as e:
                assert e.code == 6  # "Decode step failed": a target outside the embedding table
                return None
            return out.reshape(-1), st.states_1.reshape(-1), st.states_2.reshape(-1)

        n_same = n_multi = 0
        for k in range(6):
            T = 5
            enc = (0.5 * rng.standard_normal((1024, T))).astype(np.float32)
            calls.clear()
            got = oracle.greedy_decode(enc, T, step=gpu_step, single_step=False)
            ref = oracle.greedy_decode(enc, T, model, single_step=False)
            n_multi += max(calls, default=1) > 1
            if (got.rc, got.tokens, got.n_steps) == (ref.rc, ref.tokens, ref.n_steps):
                n_same += 1
            else:
                assert ref.margins.size and ref.margins.min() < NEAR_TIE, (k, got.tokens, ref.tokens)
        assert n_same >= 5 and n_multi >= 1  # at least one utterance re-fed U > 1 targets


def test_schedule_does_not_change_results(amira, monkeypatch):
    """The lane plan (M-tiles), the blank-speculation table, readiness per k-chunk and the CTA-pair form decide WHEN the tcgen05 engine computes a step,
    never its inputs: tokens, counts, step counts and final states of a mixed-length batch are bit-identical under every
    schedule of one form of the kernel (the one-CTA form sums the k-chunks in another order: same tokens, states to 1e-4) (decoder_optimized.rs:54-200 is one sequential loop per stream; scripts/decode_soak.py is the long form)."""
    rng = np.random.default_rng(77)
    B, T = 320, 96
    enc = (0.5 * rng.standard_normal((B, 1024, T))).astype(np.float32)
    lens = rng.integers(8, T + 1, B).astype(np.int64)
    ref = None
    variants = [{}, {"AMIRA_WS_TILES": "1"}, {"AMIRA_WS_TILES": "3", "AMIRA_WS_SPEC": "0"}, {"AMIRA_WS_SPEC": "3,3,3"},
                {"AMIRA_WS_SPEC": "2,2,2,2", "AMIRA_WS_TILES": "2"}, {"AMIRA_WS_CHUNK": "0"}, {"AMIRA_WS_PAIR": "0"}]
    with amira.Context(device_id=0, decode_engine=4) as c:
        c.load_weights(amira.synthetic_weights(3456))
        for v in variants:
            for k in ("AMIRA_WS_TILES", "AMIRA_WS_SPEC", "AMIRA_WS_PAIR", "AMIRA_WS_CHUNK"):
                monkeypatch.delenv(k, raising=False)
            for k, val in v.items():
                monkeypatch.setenv(k, val)
            for _ in range(2):
                toks, st, steps = c.greedy_decode(enc, lens)
                got = (toks, np.asarray(steps).tolist(), st.states_1.tobytes(), st.states_2.tobytes())
                if ref is None:
                    ref = got
                    assert all(t is not None for t in toks) and sum(len(t) for t in toks) > 0
                assert got[0] == ref[0] and got[1] == ref[1], v
                if "AMIRA_WS_PAIR" in v:  # the one-CTA form walks the k-chunks in another order: fp32 sums differ in the last bits
                    for k in (2, 3):
                        np.testing.assert_allclose(np.frombuffer(got[k], np.float32), np.frombuffer(ref[k], np.float32), rtol=1e-4, atol=1e-4)
                else:
                    assert got[2] == ref[2] and got[3] == ref[3], v
