"""CPU ORACLE for the amira-rust-asr-server hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (amira_b200 / libamira_b200.so) never does.

Two layers:
  * ctypes bindings over oracle/libamira_oracle.so (plain-C restatement, see amira_oracle.h);
  * an independent float64 numpy restatement of the mel front end (`preprocess_numpy`) and of the host-side
    string ops (`Vocabulary`), used to cross-check the C code.

Parity status: loop / layout / argmax / conversion are pinned by the reference's own KATs
(tests/test_oracle_kats.py); mel values and LSTM/joint numerics are PARITY UNPINNED — the reference's
ONNX models are absent Git-LFS pointers (see amira_oracle.h).  Citations are relative to the reference root.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libamira_oracle.so")

VOCAB, BLANK, H, ENC = 1030, 1024, 640, 1024
MAX_SYMBOLS_PER_STEP, MAX_TOTAL_TOKENS = 30, 200  # src/constants.rs:135-136
NMEL, NFFT, NBIN, WIN, HOP = 128, 512, 257, 400, 160
N_PARAMS = 8946310


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc; no GPU involved)."""
    srcs = [os.path.join(_HERE, f) for f in ("amira_oracle.c", "amira_oracle.h", "amira_oracle_frontend.inc",
                                             "../include/amira_hann400.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        cmd = [cc, "-O3", "-mavx2", "-ffp-contract=off", "-fno-math-errno", "-fopenmp", "-fPIC", "-std=c11",
               "-shared", "-o", _SO, srcs[0], "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:  # a compiler without libgomp: build single-threaded
            cmd.remove("-fopenmp")
            subprocess.run(cmd, check=True)
    return _SO


_lib = None

_STEP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int32), C.c_int,
                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int)


class _Model(C.Structure):
    _fields_ = [("emb", C.c_void_p), ("w_ih", C.c_void_p * 2), ("w_hh", C.c_void_p * 2), ("b_ih", C.c_void_p * 2),
                ("b_hh", C.c_void_p * 2), ("w_enc", C.c_void_p), ("b_enc", C.c_void_p), ("w_pred", C.c_void_p),
                ("b_pred", C.c_void_p), ("w_out", C.c_void_p), ("b_out", C.c_void_p), ("act_relu", C.c_int)]


class _Cfg(C.Structure):
    _fields_ = [("max_symbols_per_step", C.c_int), ("max_total_tokens", C.c_int), ("blank", C.c_int),
                ("single_step", C.c_int), ("state_update_on_nonblank_only", C.c_int), ("tdt_durations", C.c_int),
                ("initial_last", C.c_int)]


class _Stats(C.Structure):
    _fields_ = [("n_tokens", C.c_int), ("n_steps", C.c_int), ("frames_visited", C.c_int), ("min_margin", C.c_float)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        for name in ("orc_bytes_to_f32_optimized", "orc_bytes_to_f32_samples", "orc_bytes_to_f32_simd"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_extract_frame_into.restype = C.c_size_t
        L.orc_extract_frame_into.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                             C.c_size_t]
        L.orc_argmax_zero_copy.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_float)]
        L.orc_features_len.restype = C.c_int64
        L.orc_features_len.argtypes = [C.c_int64]
        L.orc_mel_filterbank.argtypes = [C.c_void_p]
        L.orc_hann_window_padded.argtypes = [C.c_void_p]
        L.orc_preprocess.restype = C.c_int64
        L.orc_preprocess.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]
        L.orc_preprocess_pcm16_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                                 C.c_int]
        L.orc_model_bind.argtypes = [C.POINTER(_Model), C.c_void_p]
        L.orc_model_random_init.argtypes = [C.c_void_p, C.c_uint64, C.c_float]
        L.orc_decoder_joint.argtypes = [C.POINTER(_Model), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.orc_greedy_decode.restype = C.c_int
        L.orc_greedy_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.c_void_p, _STEP_FN, C.c_void_p,
                                        C.POINTER(_Cfg), C.c_void_p, C.POINTER(_Stats), C.c_void_p, C.c_int]
        L.orc_greedy_decode_batch.restype = C.c_int
        L.orc_greedy_decode_batch.argtypes = [C.POINTER(_Model), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.POINTER(_Cfg), C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------- a1-a3: PCM conversion
def bytes_to_f32_optimized(b: bytes) -> np.ndarray:
    """src/performance_opts.rs:14-31 (odd trailing byte -> (byte as i16) / 128)."""
    src = np.frombuffer(bytes(b), dtype=np.uint8)
    out = np.empty(len(src) // 2 + 1, dtype=np.float32)
    n = lib().orc_bytes_to_f32_optimized(_p(src), len(src), _p(out))
    return out[:n].copy()


def bytes_to_f32_samples(b: bytes) -> np.ndarray:
    """src/asr/audio.rs:18-26 (odd trailing byte dropped)."""
    src = np.frombuffer(bytes(b), dtype=np.uint8)
    out = np.empty(len(src) // 2 + 1, dtype=np.float32)
    n = lib().orc_bytes_to_f32_samples(_p(src), len(src), _p(out))
    return out[:n].copy()


def bytes_to_f32_simd(b: bytes) -> np.ndarray:
    """src/asr/simd.rs:86-114,222-248."""
    src = np.frombuffer(bytes(b), dtype=np.uint8)
    out = np.empty(len(src) // 2 + 1, dtype=np.float32)
    n = lib().orc_bytes_to_f32_simd(_p(src), len(src), _p(out))
    return out[:n].copy()


# ---------------------------------------------------------------- a7 / a9
def extract_frame_into(data: np.ndarray, shape, t: int, out_len: int | None = None) -> tuple[int, np.ndarray]:
    """src/asr/zero_copy.rs:49-69."""
    data = np.ascontiguousarray(data, dtype=np.float32).ravel()
    shp = np.asarray(shape, dtype=np.uintp)
    out = np.zeros(out_len if out_len is not None else (int(shape[1]) if len(shape) == 3 else 1), dtype=np.float32)
    n = lib().orc_extract_frame_into(_p(data), data.size, _p(shp), len(shape), t, _p(out), out.size)
    return int(n), out


def argmax_zero_copy(logits) -> tuple[int, float]:
    """src/asr/zero_copy.rs:190-232."""
    a = np.ascontiguousarray(logits, dtype=np.float32).ravel()
    idx, val = C.c_size_t(0), C.c_float(0)
    lib().orc_argmax_zero_copy(_p(a), a.size, C.byref(idx), C.byref(val))
    return int(idx.value), float(val.value)


# ---------------------------------------------------------------- a4: mel front end
def features_len(n: int) -> int:
    return int(lib().orc_features_len(n))


def mel_filterbank() -> np.ndarray:
    fb = np.empty((NMEL, NBIN), dtype=np.float32)
    lib().orc_mel_filterbank(_p(fb))
    return fb


def hann_window_padded() -> np.ndarray:
    w = np.empty(NFFT, dtype=np.float64)
    lib().orc_hann_window_padded(_p(w))
    return w


def preprocess(wave: np.ndarray, precision: str = "f32", t_stride: int | None = None) -> tuple[np.ndarray, int]:
    """C oracle of the `preprocessor` model for ONE utterance ([1,N] -> [1,128,T'])."""
    wave = np.ascontiguousarray(wave, dtype=np.float32)
    L = features_len(wave.size)
    ts = t_stride if t_stride is not None else L
    out = np.zeros((NMEL, max(ts, 1)), dtype=np.float32)
    got = lib().orc_preprocess(_p(wave), wave.size, _p(out), ts, 1 if precision == "f64" else 0)
    assert got == L
    return out[:, :ts], L


def preprocess_pcm16_batch(pcm: np.ndarray, offsets: np.ndarray, t_stride: int, threads: int = 0):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    B = offsets.size - 1
    feats = np.zeros((B, NMEL, t_stride), dtype=np.float32)
    lens = np.zeros(B, dtype=np.int64)
    lib().orc_preprocess_pcm16_batch(_p(pcm), _p(offsets), B, _p(feats), t_stride, _p(lens), threads)
    return feats, lens


def _slaney_fb_numpy() -> np.ndarray:
    """librosa.filters.mel(16000, 512, 128, 0, 8000, norm='slaney') restated in numpy float64 -> float32."""
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz2mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def mel2hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0.0, 8000.0, NBIN)
    mel_f = mel2hz(np.linspace(hz2mel(0.0), hz2mel(8000.0), NMEL + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:NMEL + 2] - mel_f[:NMEL]))[:, None]
    return w.astype(np.float32)


def preprocess_numpy(wave: np.ndarray) -> tuple[np.ndarray, int]:
    """Independent float64 numpy restatement of the front-end spec, np.fft based."""
    x = np.asarray(wave, dtype=np.float32).astype(np.float64)
    n = x.size
    if n == 0:
        return np.zeros((NMEL, 0), dtype=np.float32), 0
    L = n // HOP + 1
    y = np.empty_like(x)
    y[0] = x[0]
    y[1:] = x[1:] - 0.97 * x[:-1]
    idx = np.arange(-NFFT // 2, n + NFFT // 2)
    if n > 1:
        p = 2 * (n - 1)
        idx = np.mod(idx, p)
        idx = np.where(idx < n, idx, p - idx)
    else:
        idx = np.zeros_like(idx)
    ypad = y[idx]
    win = hann_window_padded()  # float32 table of torch.hann_window(400, periodic=False), include/amira_hann400.h
    frames = np.stack([ypad[t * HOP:t * HOP + NFFT] for t in range(L)]) * win[None, :]
    power = np.abs(np.fft.rfft(frames, axis=1)) ** 2  # [L, 257]
    mel = power @ _slaney_fb_numpy().astype(np.float64).T  # [L, 128]
    logmel = np.log(mel + 2.0 ** -24)
    mean = logmel.mean(axis=0)
    std = np.sqrt(((logmel - mean) ** 2).sum(axis=0) / (L - 1)) if L > 1 else np.zeros(NMEL)
    out = (logmel - mean) / (std + 1e-5)
    return out.T.astype(np.float32), L


# ---------------------------------------------------------------- a8: model
class Model:
    """Random-init or blob-backed prediction net + joint (layout: DESIGN.md 'weight blob')."""

    def __init__(self, blob: np.ndarray | None = None, seed: int = 3456, blank_bias: float = 0.0, act: str = "tanh"):
        if blob is None:
            blob = np.empty(N_PARAMS, dtype=np.float32)
            lib().orc_model_random_init(_p(blob), seed, blank_bias)
        self.blob = np.ascontiguousarray(blob, dtype=np.float32)
        assert self.blob.size == N_PARAMS
        self._m = _Model()
        lib().orc_model_bind(C.byref(self._m), _p(self.blob))
        self._m.act_relu = 1 if act == "relu" else 0

    # views into the blob, same order as orc_model_bind
    def tensors(self) -> dict[str, np.ndarray]:
        o, b, t = 0, self.blob, {}

        def take(name, *shape):
            nonlocal o
            n = int(np.prod(shape))
            t[name] = b[o:o + n].reshape(shape)
            o += n

        take("emb", 1025, H)
        for l in range(2):
            take(f"w_ih{l}", 4 * H, H)
            take(f"w_hh{l}", 4 * H, H)
            take(f"b_ih{l}", 4 * H)
            take(f"b_hh{l}", 4 * H)
        take("w_enc", H, ENC)
        take("b_enc", H)
        take("w_pred", H, H)
        take("b_pred", H)
        take("w_out", VOCAB, H)
        take("b_out", VOCAB)
        assert o == N_PARAMS
        return t

    def decoder_joint(self, enc: np.ndarray, targets, states_1: np.ndarray, states_2: np.ndarray):
        """Triton contract op for B=1 (model-repo/decoder_joint/config.pbtxt; src/triton/model.rs:581-722).
        enc [1024,T]; targets [U]; states [2,1,640].  Returns outputs [U,T,1030], new states."""
        enc = np.ascontiguousarray(enc, dtype=np.float32).reshape(ENC, -1)
        T = enc.shape[1]
        tg = np.ascontiguousarray(targets, dtype=np.int32)
        s1 = np.ascontiguousarray(states_1, dtype=np.float32).copy()
        s2 = np.ascontiguousarray(states_2, dtype=np.float32).copy()
        out = np.empty((tg.size, T, VOCAB), dtype=np.float32)
        lib().orc_decoder_joint(C.byref(self._m), _p(enc), T, _p(tg), tg.size, _p(s1), _p(s2), _p(out))
        return out, s1, s2


@dataclass
class DecodeResult:
    tokens: list
    states_1: np.ndarray
    states_2: np.ndarray
    n_steps: int
    frames_visited: int
    min_margin: float
    margins: np.ndarray
    rc: int


def greedy_decode(enc: np.ndarray, encoded_len: int, model: Model | None = None, step=None, states=None,
                  single_step: bool = True, max_symbols: int = MAX_SYMBOLS_PER_STEP,
                  max_total: int = MAX_TOTAL_TOKENS, blank: int = BLANK, state_update_on_nonblank_only: bool = False,
                  tdt_durations: bool = False, initial_last: int | None = None) -> DecodeResult:
    """src/asr/decoder_optimized.rs:24-200 for one utterance.  `step` is an optional Python mock with the
    signature step(frame, targets, states_1, states_2) -> (logits, states_1, states_2) or None for failure."""
    enc = np.ascontiguousarray(enc, dtype=np.float32).ravel()
    s1 = np.zeros(2 * H, np.float32) if states is None else np.ascontiguousarray(states[0], np.float32).ravel().copy()
    s2 = np.zeros(2 * H, np.float32) if states is None else np.ascontiguousarray(states[1], np.float32).ravel().copy()
    cfg = _Cfg(max_symbols, max_total, blank, 1 if single_step else 0, int(state_update_on_nonblank_only), int(tdt_durations),
               blank if initial_last is None else int(initial_last))
    toks = np.zeros(max(max_total, 1) + 1, dtype=np.int32)
    st = _Stats()
    cap = (max_total + max(int(encoded_len), 0)) + 8
    margins = np.full(cap, np.inf, dtype=np.float32)
    L = lib()
    if step is None:
        fn = C.cast(L.orc_model_step, _STEP_FN)
        user = C.cast(C.pointer(model._m), C.c_void_p)
    else:
        def _cb(_user, frame, features, targets, U, ps1, ps2, logits, cap_):
            fr = np.ctypeslib.as_array(frame, shape=(features,)).copy()
            tg = np.ctypeslib.as_array(targets, shape=(U,)).copy()
            a1 = np.ctypeslib.as_array(ps1, shape=(2 * H,))
            a2 = np.ctypeslib.as_array(ps2, shape=(2 * H,))
            r = step(fr, tg, a1.copy(), a2.copy())
            if r is None:
                return -1
            lg, n1, n2 = r
            lg = np.asarray(lg, dtype=np.float32).ravel()
            if lg.size > cap_:
                return -1
            np.ctypeslib.as_array(logits, shape=(cap_,))[:lg.size] = lg
            a1[:] = np.asarray(n1, np.float32).ravel()
            a2[:] = np.asarray(n2, np.float32).ravel()
            return int(lg.size)

        fn = _STEP_FN(_cb)
        user = None
    rc = L.orc_greedy_decode(_p(enc), enc.size, int(encoded_len), _p(s1), _p(s2), fn, user, C.byref(cfg), _p(toks),
                             C.byref(st), _p(margins), cap)
    return DecodeResult(toks[:st.n_tokens].tolist(), s1.reshape(2, 1, H), s2.reshape(2, 1, H), st.n_steps,
                        st.frames_visited, float(st.min_margin), margins[:st.n_steps].copy(), rc)


def greedy_decode_batch(model: Model, enc: np.ndarray, enc_lens=None, states=None, single_step: bool = True,
                        max_symbols: int = MAX_SYMBOLS_PER_STEP, max_total: int = MAX_TOTAL_TOKENS,
                        blank: int = BLANK, threads: int = 0):
    """B independent utterances (the reference is B=1 per request); enc [B,1024,T]."""
    enc = np.ascontiguousarray(enc, dtype=np.float32)
    B, _, T = enc.shape
    lens = None if enc_lens is None else np.ascontiguousarray(enc_lens, dtype=np.int64)
    s1 = s2 = None
    if states is not None:
        s1 = np.ascontiguousarray(states[0], np.float32).copy()
        s2 = np.ascontiguousarray(states[1], np.float32).copy()
    cfg = _Cfg(max_symbols, max_total, blank, 1 if single_step else 0, 0, 0, blank)
    toks = np.zeros((B, max(max_total, 1)), dtype=np.int32)
    ntok = np.zeros(B, np.int32)
    nstep = np.zeros(B, np.int32)
    mm = np.zeros(B, np.float32)
    rc = lib().orc_greedy_decode_batch(C.byref(model._m), _p(enc), B, T, _p(lens), _p(s1), _p(s2), C.byref(cfg),
                                       _p(toks), _p(ntok), _p(nstep), _p(mm), threads)
    return dict(tokens=toks, n_tokens=ntok, n_steps=nstep, min_margin=mm, states_1=s1, states_2=s2, rc=rc)


# ---------------------------------------------------------------- a12: host-side string ops
class Vocabulary:
    """src/asr/types.rs:77-155."""

    def __init__(self, id_to_token: dict):
        self.id_to_token = dict(id_to_token)

    @classmethod
    def load_from_file(cls, path: str) -> "Vocabulary":  # types.rs:87-108
        m = {}
        with open(path, "r", encoding="utf-8") as f:
            for line in f.read().splitlines():
                parts = line.split()
                if len(parts) >= 2:
                    try:
                        m[int(parts[-1])] = " ".join(parts[:-1])
                    except ValueError:
                        pass
        return cls(m)

    def decode_tokens(self, ids) -> str:  # types.rs:111-135
        out = ""
        for i in ids:
            tok = self.id_to_token.get(int(i))
            if tok is None:
                continue  # unknown ids silently skipped (:115-116)
            out += (" " + tok[1:]) if tok.startswith("▁") else tok
        return out.strip()
