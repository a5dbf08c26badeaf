"""Python face of the C++ host pipeline (csrc/host_pipeline.cpp): mirror of the reference's `AsrPipeline` trait
(src/asr/pipeline.rs:20-67), `Transcription` (src/asr/types.rs:217-232) and `Vocabulary` (src/asr/types.rs:77-155).
All work happens behind the C ABI; the encoder model is out of scope and is injected as a callback."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from ._lib import AmiraError, Context, DecoderState, STATE_SIZE, load_library

_ENCODER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.POINTER(C.c_float)),
                          C.POINTER(C.c_int64))


class _Transcription(C.Structure):
    _fields_ = [("audio_length_samples", C.c_int64), ("features_length", C.c_int64), ("encoded_length", C.c_int64),
                ("n_tokens", C.c_int32), ("text_len", C.c_int32)]


@dataclass
class Transcription:
    """src/asr/types.rs:217-232"""
    text: str = ""
    tokens: list = field(default_factory=list)
    audio_length_samples: int = 0
    features_length: int = 0
    encoded_length: int = 0


def _bind(L):
    vp, i32 = C.c_void_p, C.c_int32
    if getattr(L, "_pipeline_bound", False):
        return
    L.amira_pipeline_create.argtypes = [vp, C.c_char_p, _ENCODER_FN, vp, C.POINTER(vp)]
    L.amira_pipeline_destroy.argtypes = [vp]
    L.amira_pipeline_last_error.argtypes = [vp]
    L.amira_pipeline_last_error.restype = C.c_char_p
    tail = [C.POINTER(_Transcription), vp, i32, vp, C.c_size_t]
    L.amira_pipeline_process_batch.argtypes = [vp, vp, C.c_size_t] + tail
    L.amira_pipeline_process_stream_chunk.argtypes = [vp, vp, C.c_size_t, vp, vp] + tail
    L.amira_pipeline_process_batch_samples.argtypes = [vp, vp, C.c_size_t] + tail
    L.amira_pipeline_process_stream_samples.argtypes = [vp, vp, C.c_size_t, vp, vp] + tail
    L.amira_vocab_decode.argtypes = [vp, vp, i32, vp, C.c_size_t, C.POINTER(i32)]
    L.amira_shard_utterances.argtypes = [vp, i32, i32, vp]
    L.amira_batcher_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.amira_batcher_destroy.argtypes = [vp]
    L.amira_batcher_process_batch.argtypes = [vp, vp, C.c_size_t] + tail
    L.amira_batcher_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L._pipeline_bound = True


def shard_utterances(costs, n_shards: int) -> np.ndarray:
    """Host-side data-parallel partition of independent utterances over GPUs (no collective)."""
    L = load_library()
    _bind(L)
    costs = np.ascontiguousarray(costs, dtype=np.int64)
    out = np.zeros(costs.size, dtype=np.int32)
    rc = L.amira_shard_utterances(costs.ctypes.data if costs.size else None, costs.size, n_shards,
                                  out.ctypes.data if costs.size else None)
    if rc:
        raise AmiraError(rc, "amira_shard_utterances")
    return out


class NativeEncoder:
    """An encoder that is already a C function with the `amira_encoder_fn` signature (what the Rust server would install): the
    address of the function and its user pointer go to the pipeline as they are, no Python frame per call."""

    def __init__(self, fn_address: int, user_address: int | None = None, keepalive=None):
        self.fn = C.cast(fn_address, _ENCODER_FN)
        self.user = C.c_void_p(user_address)
        self.keepalive = keepalive  # whatever owns the code and the user object


class B200AsrPipeline:
    """impl AsrPipeline (src/asr/pipeline.rs:20-67) backed by libamira_b200.so.

    encoder(features [128, L] f32) -> encoder outputs [1024, T] f32; stays whatever the deployment uses (a Python callable, or a
    NativeEncoder wrapping a C function)."""

    def __init__(self, ctx: Context, vocab_path: str | None, encoder):
        self._L = load_library()
        _bind(self._L)
        self._ctx = ctx
        self._encoder = encoder
        self._keep = None
        if isinstance(encoder, NativeEncoder):
            self._cb = encoder.fn
            self._h = C.c_void_p()
            rc = self._L.amira_pipeline_create(ctx.handle, os.fsencode(vocab_path) if vocab_path else None, self._cb, encoder.user,
                                               C.byref(self._h))
            if rc:
                raise AmiraError(rc, (self._L.amira_pipeline_last_error(None) or b"").decode())
            return

        def _cb(_user, feats, flen, out_ptr, out_len):
            try:
                f = np.ctypeslib.as_array(feats, shape=(128, max(int(flen), 1)))[:, :int(flen)]
                enc = np.ascontiguousarray(self._encoder(f), dtype=np.float32)
                self._keep = enc
                out_ptr[0] = enc.ctypes.data_as(C.POINTER(C.c_float))
                out_len[0] = enc.shape[1] if enc.ndim == 2 else 0
                return 0
            except Exception:  # a failing encoder fails the request, never the process
                return 1

        self._cb = _ENCODER_FN(_cb)
        self._h = C.c_void_p()
        rc = self._L.amira_pipeline_create(ctx.handle, os.fsencode(vocab_path) if vocab_path else None, self._cb, None,
                                           C.byref(self._h))
        if rc:
            raise AmiraError(rc, (self._L.amira_pipeline_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.amira_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _finish(self, rc, t, toks, text):
        if rc:
            raise AmiraError(rc, (self._L.amira_pipeline_last_error(self._h) or b"").decode())
        return Transcription(text.value.decode("utf-8", "replace"), toks[:t.n_tokens].tolist(), t.audio_length_samples,
                             t.features_length, t.encoded_length)

    def _bufs(self):
        return _Transcription(), np.zeros(self._ctx.max_total_tokens, np.int32), C.create_string_buffer(1 << 16)

    def process_batch(self, audio_bytes: bytes) -> Transcription:
        t, toks, text = self._bufs()
        b = np.frombuffer(bytes(audio_bytes), dtype=np.uint8)
        rc = self._L.amira_pipeline_process_batch(self._h, b.ctypes.data if b.size else None, b.size, C.byref(t),
                                                  toks.ctypes.data, toks.size, text, len(text))
        return self._finish(rc, t, toks, text)

    def process_stream_chunk(self, audio_bytes: bytes, state: DecoderState) -> Transcription:
        t, toks, text = self._bufs()
        b = np.frombuffer(bytes(audio_bytes), dtype=np.uint8)
        assert state.states_1.shape == (2, 1, STATE_SIZE)
        rc = self._L.amira_pipeline_process_stream_chunk(self._h, b.ctypes.data if b.size else None, b.size,
                                                         state.states_1.ctypes.data, state.states_2.ctypes.data,
                                                         C.byref(t), toks.ctypes.data, toks.size, text, len(text))
        return self._finish(rc, t, toks, text)

    def process_batch_samples(self, samples) -> Transcription:
        t, toks, text = self._bufs()
        s = np.ascontiguousarray(samples, dtype=np.float32)
        rc = self._L.amira_pipeline_process_batch_samples(self._h, s.ctypes.data if s.size else None, s.size, C.byref(t),
                                                          toks.ctypes.data, toks.size, text, len(text))
        return self._finish(rc, t, toks, text)

    def process_stream_samples(self, samples, state: DecoderState) -> Transcription:
        t, toks, text = self._bufs()
        s = np.ascontiguousarray(samples, dtype=np.float32)
        rc = self._L.amira_pipeline_process_stream_samples(self._h, s.ctypes.data if s.size else None, s.size,
                                                           state.states_1.ctypes.data, state.states_2.ctypes.data,
                                                           C.byref(t), toks.ctypes.data, toks.size, text, len(text))
        return self._finish(rc, t, toks, text)

    def decode_tokens(self, ids) -> str:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        text = C.create_string_buffer(1 << 16)
        n = C.c_int32(0)
        rc = self._L.amira_vocab_decode(self._h, ids.ctypes.data if ids.size else None, ids.size, text, len(text), C.byref(n))
        if rc:
            raise AmiraError(rc, (self._L.amira_pipeline_last_error(self._h) or b"").decode())
        return text.value.decode("utf-8", "replace")


class Batcher:
    """Request micro-batcher above a B200AsrPipeline (csrc/host_pipeline.cpp): concurrent process_batch calls from many
    threads are coalesced into one front-end launch and one decode launch; each caller gets its own Transcription."""

    def __init__(self, pipeline: B200AsrPipeline, max_batch: int = 64, max_wait_us: int = 200):
        self._p = pipeline
        self._L = pipeline._L
        self._h = C.c_void_p()
        rc = self._L.amira_batcher_create(pipeline._h, max_batch, max_wait_us, C.byref(self._h))
        if rc:
            raise AmiraError(rc, (self._L.amira_pipeline_last_error(pipeline._h) or b"").decode())

    def process_batch(self, audio_bytes: bytes) -> Transcription:
        t, toks, text = self._p._bufs()
        b = np.frombuffer(bytes(audio_bytes), dtype=np.uint8)
        rc = self._L.amira_batcher_process_batch(self._h, b.ctypes.data if b.size else None, b.size, C.byref(t),
                                                 toks.ctypes.data, toks.size, text, len(text))
        return self._p._finish(rc, t, toks, text)

    def stats(self) -> tuple[int, int]:
        """(requests that went through the queue, launches they were coalesced into)"""
        a, b = C.c_int64(0), C.c_int64(0)
        self._L.amira_batcher_stats(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.amira_batcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Vocabulary:
    """Vocabulary (src/asr/types.rs:77-155) through the C++ implementation; needs no GPU."""

    def __init__(self, path: str):
        self._L = load_library()
        _bind(self._L)
        self.path = path
        self._h = C.c_void_p()
        rc = self._L.amira_pipeline_create(None, os.fsencode(path), _ENCODER_FN(), None, C.byref(self._h))
        if rc:
            raise AmiraError(rc, (self._L.amira_pipeline_last_error(None) or b"").decode())

    @classmethod
    def load_from_file(cls, path: str) -> "Vocabulary":
        return cls(path)

    def decode_tokens(self, ids) -> str:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n = C.c_int32(0)
        text = C.create_string_buffer(1 << 16)
        rc = self._L.amira_vocab_decode(self._h, ids.ctypes.data if ids.size else None, ids.size, text, len(text), C.byref(n))
        if rc:
            raise AmiraError(rc, "amira_vocab_decode")
        return text.value.decode("utf-8", "replace")

    def __del__(self):
        try:
            if self._h.value:
                self._L.amira_pipeline_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass
