"""ctypes binding of libamira_b200.so (include/amira_b200.h is the single source of truth)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, os.environ.get("AMIRA_B200_LIB", "libamira_b200.so"))  # the override exists for A/B timing of builds

VOCAB_SIZE, BLANK_ID, STATE_SIZE, ENC_DIM, N_MELS = 1030, 1024, 640, 1024, 128
MAX_SYMBOLS_PER_STEP, MAX_TOTAL_TOKENS = 30, 200  # src/constants.rs:135-136
AMIRA_N_PARAMS = 8946310

STATUS = {0: "AMIRA_OK", 1: "AMIRA_ERR_INVALID_VALUE", 2: "AMIRA_ERR_OUT_OF_MEMORY", 3: "AMIRA_ERR_UNKNOWN",
          4: "AMIRA_ERR_NOT_READY", 5: "AMIRA_ERR_NO_DEVICE", 6: "AMIRA_ERR_DECODE_STEP", 7: "AMIRA_ERR_NO_WEIGHTS",
          8: "AMIRA_ERR_IO"}

# every symbol include/amira_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTS = ["amira_device_count", "amira_config_default", "amira_ctx_create", "amira_ctx_destroy", "amira_last_error",
           "amira_ctx_set_stream", "amira_ctx_synchronize", "amira_ctx_launch_count", "amira_ctx_profile",
           "amira_ctx_kernel_ms", "amira_debug_tc_gemm", "amira_debug_ws_trace", "amira_weights_random_init",
           "amira_ctx_load_weights", "amira_ctx_load_weights_file", "amira_features_len", "amira_preprocess_pcm16",
           "amira_preprocess_f32", "amira_bytes_to_f32", "amira_decoder_joint", "amira_greedy_decode",
           "amira_stream_open", "amira_stream_close", "amira_stream_get_state", "amira_stream_set_state",
           "amira_stream_decode", "amira_pipeline_create", "amira_pipeline_destroy", "amira_pipeline_process_batch",
           "amira_pipeline_process_stream_chunk", "amira_pipeline_process_batch_samples",
           "amira_pipeline_process_stream_samples", "amira_pipeline_last_error", "amira_vocab_decode",
           "amira_shard_utterances", "amira_batcher_create", "amira_batcher_destroy", "amira_batcher_process_batch",
           "amira_batcher_stats", "amira_ctx_max_total_tokens", "amira_preprocess_pcm16_packed",
           "amira_greedy_decode_packed", "amira_preprocess_f32_packed", "amira_weave_transcript_segs", "amira_best_alignment",
           "amira_is_overlap_silence", "amira_mean_amplitude", "amira_window_sequence", "amira_stream_group_create",
           "amira_stream_group_destroy", "amira_stream_group_last_error", "amira_stream_group_clear",
           "amira_stream_group_process_chunks", "amira_stream_group_transcript", "amira_stream_group_tokens",
           "amira_stream_group_audio_length", "amira_stream_group_process_batch", "amira_stream_group_stats", "amira_ctx_fork", "amira_device_alloc",
           "amira_device_free", "amira_ipc_export", "amira_ipc_import", "amira_ipc_close", "amira_wire_classify_frame",
           "amira_wire_parse_batch_request", "amira_wire_format_response", "amira_device_reset", "amira_logmel_pcm16_packed",
           "amira_greedy_decode_resume", "amira_stream_group_set_incremental", "amira_stream_group_flush", "amira_stream_group_progress"]


class AmiraError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS.get(code, code)}: {message}")
        self.code = code
        self.message = message


class _Config(C.Structure):
    _fields_ = [("device_id", C.c_int32), ("max_symbols_per_step", C.c_int32), ("max_total_tokens", C.c_int32),
                ("blank_id", C.c_int32), ("joint_activation", C.c_int32), ("decode_engine", C.c_int32),
                ("max_streams", C.c_int32), ("decode_rule", C.c_int32)]


_lib = None


def lib_path() -> str:
    return _SO


def load_library():
    """Loads libamira_b200.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). amira_b200 has no CPU fallback.")
    L = C.CDLL(_SO)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.amira_device_count.argtypes = [C.POINTER(i32)]
    L.amira_config_default.argtypes = [C.POINTER(_Config)]
    L.amira_device_reset.argtypes = [i32]
    L.amira_ctx_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    L.amira_ctx_destroy.argtypes = [vp]
    L.amira_ctx_fork.argtypes = [vp, C.POINTER(vp)]
    L.amira_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.amira_device_free.argtypes = [vp, vp]
    L.amira_ipc_export.argtypes = [vp, vp, vp]
    L.amira_ipc_import.argtypes = [vp, vp, C.POINTER(vp)]
    L.amira_ipc_close.argtypes = [vp, vp]
    L.amira_wire_classify_frame.argtypes = [vp, C.c_size_t, i32, C.POINTER(i32)]
    L.amira_wire_parse_batch_request.argtypes = [C.c_char_p, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                                 C.POINTER(C.c_size_t), vp, C.c_size_t]
    L.amira_wire_format_response.argtypes = [C.c_char_p, i32, C.c_char_p, vp, vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.amira_last_error.argtypes = [vp]
    L.amira_last_error.restype = C.c_char_p
    L.amira_ctx_set_stream.argtypes = [vp, vp]
    L.amira_ctx_synchronize.argtypes = [vp]
    L.amira_ctx_launch_count.argtypes = [vp, C.POINTER(i64)]
    L.amira_ctx_max_total_tokens.argtypes = [vp, C.POINTER(i32)]
    L.amira_ctx_profile.argtypes = [vp, i32]
    L.amira_ctx_kernel_ms.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(i64)]
    L.amira_debug_tc_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    L.amira_debug_ws_trace.argtypes = [vp, vp, i32]
    L.amira_weights_random_init.argtypes = [vp, C.c_size_t, C.c_uint64, C.c_float]
    L.amira_ctx_load_weights.argtypes = [vp, vp, C.c_size_t]
    L.amira_ctx_load_weights_file.argtypes = [vp, C.c_char_p]
    L.amira_features_len.argtypes = [i64, C.POINTER(i64)]
    L.amira_preprocess_pcm16.argtypes = [vp, vp, vp, i32, vp, i64, vp]
    L.amira_preprocess_f32.argtypes = [vp, vp, i64, vp, i32, vp, i64, vp]
    L.amira_preprocess_pcm16_packed.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    L.amira_logmel_pcm16_packed.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    L.amira_greedy_decode_resume.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.amira_stream_group_set_incremental.argtypes = [vp, i32]
    L.amira_stream_group_flush.argtypes = [vp, i32, vp, vp]
    L.amira_stream_group_progress.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.amira_greedy_decode_packed.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.amira_preprocess_f32_packed.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    L.amira_weave_transcript_segs.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_float, vp, C.c_size_t, C.POINTER(i32)]
    L.amira_best_alignment.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.POINTER(i32), C.POINTER(C.c_float)]
    L.amira_is_overlap_silence.argtypes = [vp, C.c_size_t, C.c_float, C.POINTER(i32)]
    L.amira_mean_amplitude.argtypes = [vp, C.c_size_t, C.POINTER(C.c_float)]
    L.amira_window_sequence.argtypes = [i64, i64, i64, i64, vp, vp, i32, C.POINTER(i32)]
    L.amira_stream_group_create.argtypes = [vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(vp)]
    L.amira_stream_group_destroy.argtypes = [vp]
    L.amira_stream_group_last_error.argtypes = [vp]
    L.amira_stream_group_last_error.restype = C.c_char_p
    L.amira_stream_group_clear.argtypes = [vp, i32]
    L.amira_stream_group_process_chunks.argtypes = [vp, i32, vp, vp, vp, vp]
    L.amira_stream_group_transcript.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(i32)]
    L.amira_stream_group_tokens.argtypes = [vp, i32, vp, i32, C.POINTER(i32)]
    L.amira_stream_group_audio_length.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.amira_stream_group_process_batch.argtypes = [vp, i32, vp, C.c_size_t, vp, vp, i32, vp, C.c_size_t]
    L.amira_stream_group_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.amira_bytes_to_f32.argtypes = [vp, vp, C.c_size_t, i32, vp, C.POINTER(C.c_size_t)]
    L.amira_decoder_joint.argtypes = [vp, vp, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.amira_greedy_decode.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.amira_stream_open.argtypes = [vp, C.POINTER(i32)]
    L.amira_stream_close.argtypes = [vp, i32]
    L.amira_stream_get_state.argtypes = [vp, i32, vp, vp]
    L.amira_stream_set_state.argtypes = [vp, i32, vp, vp]
    L.amira_stream_decode.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)  # AttributeError here = header/.so drift
        if name not in ("amira_last_error", "amira_pipeline_last_error", "amira_stream_group_last_error"):
            fn.restype = i32
    _lib = L
    return L


def device_count() -> int:
    n = C.c_int32(0)
    load_library().amira_device_count(C.byref(n))
    return int(n.value)


def device_reset(device_id: int = 0):
    """cudaDeviceReset after a sticky device error; every Context of that device must have been closed first."""
    rc = load_library().amira_device_reset(device_id)
    if rc:
        raise AmiraError(rc, "amira_device_reset")


def features_len(n_samples: int) -> int:
    """features_lens rule of the preprocessor model: floor(n/160)+1."""
    out = C.c_int64(0)
    load_library().amira_features_len(int(n_samples), C.byref(out))
    return int(out.value)


def random_weights(seed: int = 3456, blank_bias: float = 0.0) -> np.ndarray:
    """Seeded stand-in for the absent decoder_joint ONNX weights (host-side; flat fp32 blob)."""
    blob = np.empty(AMIRA_N_PARAMS, dtype=np.float32)
    rc = load_library().amira_weights_random_init(blob.ctypes.data, blob.size, seed, blank_bias)
    if rc:
        raise AmiraError(rc, "amira_weights_random_init")
    return blob


def blob_views(blob: np.ndarray) -> dict:
    """Named views into the flat weight blob (order documented in include/amira_b200.h)."""
    H, o, t = STATE_SIZE, 0, {}

    def take(name, *shape):
        nonlocal o
        n = int(np.prod(shape))
        t[name] = blob[o:o + n].reshape(shape)
        o += n

    take("emb", 1025, H)
    for l in range(2):
        take(f"w_ih{l}", 4 * H, H)
        take(f"w_hh{l}", 4 * H, H)
        take(f"b_ih{l}", 4 * H)
        take(f"b_hh{l}", 4 * H)
    take("w_enc", H, ENC_DIM)
    take("b_enc", H)
    take("w_pred", H, H)
    take("b_pred", H)
    take("w_out", VOCAB_SIZE, H)
    take("b_out", VOCAB_SIZE)
    assert o == AMIRA_N_PARAMS
    return t


# Synthetic benchmark model (BASELINE config 3; the real weights are an absent LFS object).  Plain PyTorch-style
# random init gives a degenerate decoder (one token repeated 30x per frame or never anything), so the seeded
# blob is rescaled until the prediction net matters and the blank bias is set so that ~0.25 tokens are emitted per
# encoder frame on N(0, 0.5^2) encoder outputs.  Recipe found by bisection on the CPU oracle (DESIGN.md "workload");
# tests/test_oracle_decoder.py re-checks the realised rate.  Duration logits 1025..1029 are pushed down so the
# reference's flat 1030-way argmax never leaves the embedding table.
SYNTH_SCALE = {"w_enc": 2.0, "w_pred": 12.0, "w_out": 4.0, "emb": 6.0, "w_ih0": 2.0, "w_ih1": 2.0}
SYNTH_BLANK_BIAS = 5.25


def synthetic_weights(seed: int = 3456, blank_bias: float = SYNTH_BLANK_BIAS) -> np.ndarray:
    blob = random_weights(seed, 0.0)
    t = blob_views(blob)
    for k, v in SYNTH_SCALE.items():
        t[k] *= np.float32(v)
    t["b_out"][1025:1030] -= np.float32(100.0)
    t["b_out"][BLANK_ID] += np.float32(blank_bias)
    return blob


@dataclass
class DecoderState:
    """src/asr/types.rs:159-183 — states_1 / states_2, each [2, B, 640], zeros at start."""
    states_1: np.ndarray
    states_2: np.ndarray

    @classmethod
    def new(cls, batch: int = 1) -> "DecoderState":
        return cls(np.zeros((2, batch, STATE_SIZE), np.float32), np.zeros((2, batch, STATE_SIZE), np.float32))


def _ptr(x):
    """numpy array -> host pointer; int -> raw (device) pointer; None -> NULL."""
    if x is None:
        return None
    if isinstance(x, (int, np.integer)):
        return C.c_void_p(int(x))
    return C.c_void_p(x.ctypes.data)


class Context:
    """One per GPU (amira_ctx).  Thread-safe: calls are serialised inside the library."""

    def __init__(self, device_id: int = 0, max_symbols_per_step: int = MAX_SYMBOLS_PER_STEP,
                 max_total_tokens: int = MAX_TOTAL_TOKENS, blank_id: int = BLANK_ID, joint_activation: str = "tanh",
                 decode_engine: int = 0, max_streams: int = 1024, decode_rule: int = 0):
        self._L = load_library()
        cfg = _Config()
        self._L.amira_config_default(C.byref(cfg))
        cfg.device_id, cfg.max_symbols_per_step, cfg.max_total_tokens = device_id, max_symbols_per_step, max_total_tokens
        cfg.blank_id, cfg.joint_activation = blank_id, 1 if joint_activation == "relu" else 0
        cfg.decode_engine, cfg.max_streams, cfg.decode_rule = decode_engine, max_streams, decode_rule
        self.max_total_tokens = max_total_tokens
        self._h = C.c_void_p()
        rc = self._L.amira_ctx_create(C.byref(cfg), C.byref(self._h))
        if rc:
            self._h = C.c_void_p()
            raise AmiraError(rc, (self._L.amira_last_error(None) or b"").decode())

    # -- lifecycle
    def fork(self) -> "Context":
        """Another submission lane on the same GPU sharing this context's weights (amira_ctx_fork)."""
        lane = object.__new__(Context)
        lane._L, lane.max_total_tokens, lane._h = self._L, self.max_total_tokens, C.c_void_p()
        self._check(self._L.amira_ctx_fork(self._h, C.byref(lane._h)))
        return lane

    # -- device-resident hand-off (src/cuda/cuda_helper.cu:63-183)
    def device_alloc(self, n_bytes: int) -> int:
        p = C.c_void_p()
        self._check(self._L.amira_device_alloc(self._h, n_bytes, C.byref(p)))
        return int(p.value)

    def device_free(self, dev_ptr: int):
        self._check(self._L.amira_device_free(self._h, C.c_void_p(dev_ptr)))

    def ipc_export(self, dev_ptr: int) -> bytes:
        h = (C.c_ubyte * 64)()
        self._check(self._L.amira_ipc_export(self._h, C.c_void_p(dev_ptr), C.cast(h, C.c_void_p)))
        return bytes(h)

    def ipc_import(self, handle: bytes) -> int:
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self._L.amira_ipc_import(self._h, C.cast(h, C.c_void_p), C.byref(p)))
        return int(p.value)

    def ipc_close(self, dev_ptr: int):
        self._check(self._L.amira_ipc_close(self._h, C.c_void_p(dev_ptr)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.amira_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise AmiraError(rc, (self._L.amira_last_error(self._h) or b"").decode())

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        self._check(self._L.amira_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self._L.amira_ctx_synchronize(self._h))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        self._check(self._L.amira_ctx_launch_count(self._h, C.byref(n)))
        return int(n.value)

    KERNELS = {"fe_logmel": 0, "fe_normalize": 1, "enc_proj": 2, "greedy": 3, "bytes_to_f32": 4}

    def profile(self, enable: bool = True):
        self._check(self._L.amira_ctx_profile(self._h, int(enable)))

    def kernel_ms(self, name: str):
        """(total device ms, launches) of one library kernel since profile(True)."""
        ms, n = C.c_double(0), C.c_int64(0)
        self._check(self._L.amira_ctx_kernel_ms(self._h, self.KERNELS[name], C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def debug_tc_gemm(self, A: np.ndarray, W: np.ndarray, bias: np.ndarray | None = None) -> np.ndarray:
        """C = A W^T + bias through the tcgen05 split-bf16 GEMM (diagnostics / unit test hook)."""
        A = np.ascontiguousarray(A, np.float32)
        W = np.ascontiguousarray(W, np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, np.float32)
        C_ = np.empty((A.shape[0], W.shape[0]), np.float32)
        self._check(self._L.amira_debug_tc_gemm(self._h, _ptr(A), _ptr(W), _ptr(b), A.shape[0], W.shape[0], A.shape[1], _ptr(C_)))
        return C_

    def debug_ws_trace(self, n_its: int = 512) -> np.ndarray:
        """[n_its][8 M-tiles][32] globaltimer stamps (ns) of the last weight-stationary greedy launch (AMIRA_WS_TRACE=1)."""
        out = np.zeros((n_its, 8, 32), np.int64)
        self._check(self._L.amira_debug_ws_trace(self._h, _ptr(out), n_its))
        return out

    # -- weights
    def load_weights(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        self._check(self._L.amira_ctx_load_weights(self._h, blob.ctypes.data, blob.size))

    def load_weights_file(self, path: str):
        self._check(self._L.amira_ctx_load_weights_file(self._h, os.fsencode(path)))

    # -- stage 1
    def bytes_to_f32(self, data: bytes, drop_odd: bool = False) -> np.ndarray:
        src = np.frombuffer(bytes(data), dtype=np.uint8)
        out = np.empty(src.size // 2 + 1, dtype=np.float32)
        n = C.c_size_t(0)
        self._check(self._L.amira_bytes_to_f32(self._h, _ptr(src) if src.size else None, src.size, int(drop_odd),
                                               _ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def preprocess_pcm16(self, pcm: np.ndarray, offsets, t_stride: int | None = None):
        """B utterances packed back to back (utterance b = pcm[offsets[b]:offsets[b+1]]) -> (features [B,128,t_stride],
        features_lens [B]).  Replaces convert_audio + PreprocessorModel::infer (src/asr/pipeline.rs:127-139,283-291)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        B = offsets.size - 1
        lens = np.zeros(B, dtype=np.int64)
        if t_stride is None:
            t_stride = max([features_len(int(offsets[b + 1] - offsets[b])) for b in range(B)] + [1])
        feats = np.empty((B, N_MELS, t_stride), dtype=np.float32)
        self._check(self._L.amira_preprocess_pcm16(self._h, _ptr(pcm), _ptr(offsets), B, _ptr(feats), t_stride, _ptr(lens)))
        return feats, lens

    def preprocess_pcm16_packed(self, pcm: np.ndarray, offsets):
        """Ragged form of preprocess_pcm16: returns a list of B dense [128, features_len_b] arrays (views into one packed
        buffer) and features_lens — the per-request tensors the reference gives its encoder (src/triton/model.rs:126-141)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        B = offsets.size - 1
        lens = np.zeros(B, dtype=np.int64)
        foff = np.zeros(B + 1, dtype=np.int64)
        for b in range(B):
            foff[b + 1] = foff[b] + N_MELS * features_len(int(offsets[b + 1] - offsets[b]))
        feats = np.empty(max(int(foff[B]), 1), dtype=np.float32)
        self._check(self._L.amira_preprocess_pcm16_packed(self._h, _ptr(pcm), _ptr(offsets), B, _ptr(feats), _ptr(foff), _ptr(lens)))
        return [feats[foff[b]:foff[b + 1]].reshape(N_MELS, -1) for b in range(B)], lens

    def logmel_pcm16_packed(self, pcm: np.ndarray, offsets):
        """Un-normalised log-mel (amira_logmel_pcm16_packed): a list of B [128, features_len_b] arrays and features_lens."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        B = offsets.size - 1
        lens = np.zeros(B, dtype=np.int64)
        foff = np.zeros(B + 1, dtype=np.int64)
        for b in range(B):
            foff[b + 1] = foff[b] + N_MELS * features_len(int(offsets[b + 1] - offsets[b]))
        feats = np.empty(max(int(foff[B]), 1), dtype=np.float32)
        self._check(self._L.amira_logmel_pcm16_packed(self._h, _ptr(pcm), _ptr(offsets), B, _ptr(feats), _ptr(foff), _ptr(lens)))
        return [feats[foff[b]:foff[b + 1]].reshape(N_MELS, -1) for b in range(B)], lens

    def greedy_decode_resume(self, encoder_outputs: list, state: "DecoderState", last_tokens):
        """amira_greedy_decode_resume: like greedy_decode_packed, with the last emitted token of every stream carried in and out.
        Returns (tokens per stream, DecoderState, last_tokens [B], n_steps)."""
        B = len(encoder_outputs)
        lens = np.array([int(e.shape[-1]) for e in encoder_outputs], dtype=np.int64)
        eoff = np.zeros(B + 1, dtype=np.int64)
        eoff[1:] = np.cumsum(lens * ENC_DIM)
        enc = np.empty(max(int(eoff[B]), 1), dtype=np.float32)
        for b, e in enumerate(encoder_outputs):
            enc[eoff[b]:eoff[b + 1]] = np.ascontiguousarray(e, dtype=np.float32).reshape(-1)
        st = DecoderState(np.ascontiguousarray(state.states_1, np.float32).copy(), np.ascontiguousarray(state.states_2, np.float32).copy())
        last = np.ascontiguousarray(last_tokens, dtype=np.int32).copy()
        toks = np.zeros((B, self.max_total_tokens), dtype=np.int32)
        ntok = np.zeros(B, dtype=np.int32)
        nsteps = np.zeros(B, dtype=np.int32)
        self._check(self._L.amira_greedy_decode_resume(self._h, _ptr(enc), _ptr(eoff), B, _ptr(lens), _ptr(st.states_1), _ptr(st.states_2),
                                                       _ptr(last), _ptr(toks), _ptr(ntok), _ptr(nsteps)))
        return [toks[b, :max(int(ntok[b]), 0)].tolist() for b in range(B)], st, last, nsteps

    def preprocess_pcm16_packed_raw(self, pcm_ptr: int, offsets: np.ndarray, B: int, features_ptr: int, feat_offsets: np.ndarray,
                                    lens_out: np.ndarray):
        self._check(self._L.amira_preprocess_pcm16_packed(self._h, _ptr(pcm_ptr), _ptr(offsets), B, _ptr(features_ptr),
                                                          _ptr(feat_offsets), _ptr(lens_out)))

    def logmel_pcm16_packed_raw(self, pcm_ptr: int, offsets: np.ndarray, B: int, features_ptr: int, feat_offsets: np.ndarray,
                                lens_out: np.ndarray):
        self._check(self._L.amira_logmel_pcm16_packed(self._h, _ptr(pcm_ptr), _ptr(offsets), B, _ptr(features_ptr),
                                                      _ptr(feat_offsets), _ptr(lens_out)))

    def greedy_decode_packed_raw(self, enc_ptr: int, enc_offsets: np.ndarray, B: int, lens: np.ndarray, tokens_ptr: int, ntok_ptr: int,
                                 nsteps_ptr: int | None = None, s1_ptr: int | None = None, s2_ptr: int | None = None):
        self._check(self._L.amira_greedy_decode_packed(self._h, _ptr(enc_ptr), _ptr(enc_offsets), B, _ptr(lens), _ptr(s1_ptr),
                                                       _ptr(s2_ptr), _ptr(tokens_ptr), _ptr(ntok_ptr), _ptr(nsteps_ptr)))

    def greedy_decode_packed(self, encoder_outputs: list, state: DecoderState | None = None, allow_failed: bool = False):
        """Ragged form of greedy_decode: encoder_outputs is a list of B arrays [1024, T_b] (the per-request encoder tensors of
        src/triton/model.rs:298-420); only valid frames are uploaded.  Same return value as greedy_decode."""
        B = len(encoder_outputs)
        lens = np.array([int(e.shape[-1]) for e in encoder_outputs], dtype=np.int64)
        eoff = np.zeros(B + 1, dtype=np.int64)
        eoff[1:] = np.cumsum(lens * ENC_DIM)
        enc = np.empty(max(int(eoff[B]), 1), dtype=np.float32)
        for b, e in enumerate(encoder_outputs):
            enc[eoff[b]:eoff[b + 1]] = np.ascontiguousarray(e, dtype=np.float32).reshape(-1)
        st = DecoderState.new(B) if state is None else DecoderState(
            np.ascontiguousarray(state.states_1, np.float32).copy(), np.ascontiguousarray(state.states_2, np.float32).copy())
        toks = np.zeros((B, self.max_total_tokens), dtype=np.int32)
        ntok = np.zeros(B, dtype=np.int32)
        nsteps = np.zeros(B, dtype=np.int32)
        rc = self._L.amira_greedy_decode_packed(self._h, _ptr(enc), _ptr(eoff), B, _ptr(lens), _ptr(st.states_1), _ptr(st.states_2),
                                                _ptr(toks), _ptr(ntok), _ptr(nsteps))
        if rc and not (allow_failed and rc == 6):
            self._check(rc)
        return [toks[b, :max(int(ntok[b]), 0)].tolist() if ntok[b] >= 0 else None for b in range(B)], st, nsteps

    def preprocessor(self, waveforms: np.ndarray, waveforms_lens, t_stride: int | None = None):
        """Triton contract form (model-repo/preprocessor/config.pbtxt): waveforms [B,N] f32, waveforms_lens [B] i64
        -> features [B,128,T'], features_lens [B]."""
        waveforms = np.ascontiguousarray(waveforms, dtype=np.float32)
        if waveforms.ndim == 1:
            waveforms = waveforms[None, :]
        B, N = waveforms.shape
        wl = np.ascontiguousarray(waveforms_lens, dtype=np.int64).reshape(B)
        lens = np.zeros(B, dtype=np.int64)
        if t_stride is None:
            t_stride = max([features_len(int(x)) for x in wl] + [1])
        feats = np.empty((B, N_MELS, t_stride), dtype=np.float32)
        self._check(self._L.amira_preprocess_f32(self._h, _ptr(waveforms), N, _ptr(wl), B, _ptr(feats), t_stride, _ptr(lens)))
        return feats, lens

    # raw-pointer forms (device or host pointers as ints) for callers that own device memory (bench.py)
    def preprocess_pcm16_raw(self, pcm_ptr: int, offsets: np.ndarray, B: int, features_ptr: int, t_stride: int,
                             lens_out: np.ndarray):
        self._check(self._L.amira_preprocess_pcm16(self._h, _ptr(pcm_ptr), _ptr(offsets), B, _ptr(features_ptr), t_stride,
                                                   _ptr(lens_out)))

    def greedy_decode_raw(self, enc_ptr: int, B: int, T: int, lens: np.ndarray | None, tokens_ptr: int, ntok_ptr: int,
                          nsteps_ptr: int | None = None, s1_ptr: int | None = None, s2_ptr: int | None = None):
        self._check(self._L.amira_greedy_decode(self._h, _ptr(enc_ptr), B, T, _ptr(lens), _ptr(s1_ptr), _ptr(s2_ptr),
                                                _ptr(tokens_ptr), _ptr(ntok_ptr), _ptr(nsteps_ptr)))

    # -- stage 2
    def decoder_joint(self, encoder_outputs: np.ndarray, targets: np.ndarray, target_length=None,
                      state: DecoderState | None = None):
        """Triton contract op (model-repo/decoder_joint/config.pbtxt): encoder_outputs [B,1024,T], targets [B,U] ->
        outputs [B,U,T,1030], prednet_lengths [B], DecoderState."""
        enc = np.ascontiguousarray(encoder_outputs, dtype=np.float32)
        B, _, T = enc.shape
        tg = np.ascontiguousarray(targets, dtype=np.int32).reshape(B, -1)
        U = tg.shape[1]
        tl = None if target_length is None else np.ascontiguousarray(target_length, dtype=np.int32).reshape(B)
        s1 = None if state is None else np.ascontiguousarray(state.states_1, dtype=np.float32)
        s2 = None if state is None else np.ascontiguousarray(state.states_2, dtype=np.float32)
        out = np.empty((B, U, T, VOCAB_SIZE), dtype=np.float32)
        pl = np.zeros(B, dtype=np.int32)
        o1 = np.empty((2, B, STATE_SIZE), np.float32)
        o2 = np.empty((2, B, STATE_SIZE), np.float32)
        self._check(self._L.amira_decoder_joint(self._h, _ptr(enc), B, T, _ptr(tg), U, _ptr(tl), _ptr(s1), _ptr(s2),
                                                _ptr(out), _ptr(pl), _ptr(o1), _ptr(o2)))
        return out, pl, DecoderState(o1, o2)

    def greedy_decode(self, encoder_outputs: np.ndarray, encoded_lengths=None, state: DecoderState | None = None,
                      allow_failed: bool = False):
        """B independent utterances through the persistent decode kernel.  Returns (tokens list per utterance,
        DecoderState, n_steps [B]).  Replaces greedy_decode + the per-step RPC (src/asr/decoder_optimized.rs:24-200)."""
        enc = np.ascontiguousarray(encoder_outputs, dtype=np.float32)
        B, _, T = enc.shape
        lens = None if encoded_lengths is None else np.ascontiguousarray(encoded_lengths, dtype=np.int64).reshape(B)
        st = DecoderState.new(B) if state is None else DecoderState(
            np.ascontiguousarray(state.states_1, np.float32).copy(), np.ascontiguousarray(state.states_2, np.float32).copy())
        toks = np.zeros((B, self.max_total_tokens), dtype=np.int32)
        ntok = np.zeros(B, dtype=np.int32)
        nsteps = np.zeros(B, dtype=np.int32)
        rc = self._L.amira_greedy_decode(self._h, _ptr(enc), B, T, _ptr(lens), _ptr(st.states_1), _ptr(st.states_2),
                                         _ptr(toks), _ptr(ntok), _ptr(nsteps))
        if rc and not (allow_failed and rc == 6):
            self._check(rc)
        return [toks[b, :max(int(ntok[b]), 0)].tolist() if ntok[b] >= 0 else None for b in range(B)], st, nsteps

    # -- WebSocket path: device-resident stream state
    def stream_open(self) -> int:
        s = C.c_int32(-1)
        self._check(self._L.amira_stream_open(self._h, C.byref(s)))
        return int(s.value)

    def stream_close(self, slot: int):
        self._check(self._L.amira_stream_close(self._h, slot))

    def stream_get_state(self, slot: int) -> DecoderState:
        st = DecoderState.new(1)
        self._check(self._L.amira_stream_get_state(self._h, slot, _ptr(st.states_1), _ptr(st.states_2)))
        return st

    def stream_set_state(self, slot: int, state: DecoderState):
        s1 = np.ascontiguousarray(state.states_1, np.float32)
        s2 = np.ascontiguousarray(state.states_2, np.float32)
        self._check(self._L.amira_stream_set_state(self._h, slot, _ptr(s1), _ptr(s2)))

    def stream_decode_raw(self, slots: np.ndarray, enc_ptr: int, T: int, lens: np.ndarray | None, tokens_ptr: int, ntok_ptr: int,
                          nsteps_ptr: int | None = None):
        """One tick on device-resident state slots with caller-managed (host or device) buffers."""
        self._check(self._L.amira_stream_decode(self._h, _ptr(slots), int(slots.size), _ptr(enc_ptr), T, _ptr(lens),
                                                _ptr(tokens_ptr), _ptr(ntok_ptr), _ptr(nsteps_ptr)))

    def stream_decode(self, slots, encoder_outputs: np.ndarray, encoded_lengths=None):
        enc = np.ascontiguousarray(encoder_outputs, dtype=np.float32)
        n, _, T = enc.shape
        sl = np.ascontiguousarray(slots, dtype=np.int32).reshape(n)
        lens = None if encoded_lengths is None else np.ascontiguousarray(encoded_lengths, dtype=np.int64).reshape(n)
        toks = np.zeros((n, self.max_total_tokens), dtype=np.int32)
        ntok = np.zeros(n, dtype=np.int32)
        nsteps = np.zeros(n, dtype=np.int32)
        self._check(self._L.amira_stream_decode(self._h, _ptr(sl), n, _ptr(enc), T, _ptr(lens), _ptr(toks), _ptr(ntok),
                                                _ptr(nsteps)))
        return [toks[b, :ntok[b]].tolist() for b in range(n)], nsteps
