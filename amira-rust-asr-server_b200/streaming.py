"""Python face of the streaming orchestrator (csrc/host_stream.cpp): mirror of the reference's `IncrementalAsr`
(src/asr/incremental.rs:35-298), `window_sequence` / `OverlappingAudioBuffer` (src/asr/audio.rs:72-293) and transcript
weaving (src/asr/weaving.rs).  All work happens behind the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import AmiraError, load_library
from .pipeline import B200AsrPipeline, Transcription, _Transcription

# the parameters the reference's WebSocket handler uses (src/server/stream.rs:106-109)
CHUNK_SIZE, LEADING_CONTEXT, TRAILING_CONTEXT, BUFFER_CAPACITY_SECONDS = 2.0, 1.0, 0.5, 10.0
MIN_ALIGNMENT_SCORE = 0.01  # src/asr/incremental.rs:19


def _check(rc: int, what: str):
    if rc:
        raise AmiraError(rc, what)


def weave_transcript_segs(first_seg: str, second_seg: str, percent_time_overlap: float,
                          min_alignment_score: float = MIN_ALIGNMENT_SCORE) -> str:
    """src/asr/weaving.rs:209-280"""
    L = load_library()
    cap = 4 * (len(first_seg.encode()) + len(second_seg.encode())) + 16
    out = C.create_string_buffer(cap)
    n = C.c_int32(0)
    _check(L.amira_weave_transcript_segs(first_seg.encode(), second_seg.encode(), percent_time_overlap, min_alignment_score,
                                         out, cap, C.byref(n)), "weave_transcript_segs")
    return out.raw[:n.value].decode("utf-8")


def best_alignment(first: str, second: str, percent_time_overlap: float):
    """src/asr/weaving.rs:180-203 -> (overlap in chars, score)"""
    L = load_library()
    o, s = C.c_int32(0), C.c_float(0)
    _check(L.amira_best_alignment(first.encode(), second.encode(), percent_time_overlap, C.byref(o), C.byref(s)), "best_alignment")
    return int(o.value), float(s.value)


def is_overlap_silence(overlap_audio, mean_amplitude: float) -> bool:
    """src/asr/weaving.rs:285-313"""
    L = load_library()
    a = np.ascontiguousarray(overlap_audio, dtype=np.float32)
    r = C.c_int32(0)
    _check(L.amira_is_overlap_silence(a.ctypes.data if a.size else None, a.size, mean_amplitude, C.byref(r)), "is_overlap_silence")
    return bool(r.value)


def mean_amplitude(samples) -> float:
    """src/performance_opts.rs:35-60"""
    L = load_library()
    a = np.ascontiguousarray(samples, dtype=np.float32)
    m = C.c_float(0)
    _check(L.amira_mean_amplitude(a.ctypes.data if a.size else None, a.size, C.byref(m)), "mean_amplitude")
    return float(m.value)


def window_sequence(total_len: int, window_size: int, leading_context: int, trailing_context: int):
    """src/asr/audio.rs:72-132 -> list of ((source start, end), (target start, end), overlap ratio)"""
    L = load_library()
    n = C.c_int32(0)
    _check(L.amira_window_sequence(total_len, window_size, leading_context, trailing_context, None, None, 0, C.byref(n)), "window_sequence")
    sl = np.zeros((max(n.value, 1), 4), np.int64)
    ov = np.zeros(max(n.value, 1), np.float32)
    _check(L.amira_window_sequence(total_len, window_size, leading_context, trailing_context, sl.ctypes.data, ov.ctypes.data,
                                   n.value, C.byref(n)), "window_sequence")
    return [((int(sl[i, 0]), int(sl[i, 1])), (int(sl[i, 2]), int(sl[i, 3])), np.float32(ov[i])) for i in range(n.value)]


class StreamGroup:
    """n_streams IncrementalAsr objects over one pipeline; process_chunks advances many streams by one chunk with one
    batched front-end launch and one decode launch per window round (csrc/host_stream.cpp)."""

    def __init__(self, pipeline: B200AsrPipeline, n_streams: int, chunk_size: float = CHUNK_SIZE,
                 leading_context: float = LEADING_CONTEXT, trailing_context: float = TRAILING_CONTEXT,
                 buffer_capacity: float = BUFFER_CAPACITY_SECONDS):
        self._p = pipeline
        self._L = load_library()
        self.n_streams = n_streams
        self._h = C.c_void_p()
        _check(self._L.amira_stream_group_create(pipeline._h, n_streams, chunk_size, leading_context, trailing_context,
                                                 buffer_capacity, C.byref(self._h)), "stream group create")

    def _err(self) -> str:
        return (self._L.amira_stream_group_last_error(self._h) or b"").decode()

    def process_chunks(self, streams, chunks, raise_on_error: bool = True):
        """IncrementalAsr::process_chunk (src/asr/incremental.rs:111-129) for several distinct streams at once.
        Returns the current transcript of each of them."""
        n = len(streams)
        ids = np.ascontiguousarray(streams, dtype=np.int32)
        bufs = [np.frombuffer(bytes(c), dtype=np.uint8) for c in chunks]
        ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data if b.size else None for b in bufs])
        lens = (C.c_size_t * max(n, 1))(*[b.size for b in bufs])
        status = np.zeros(max(n, 1), np.int32)
        rc = self._L.amira_stream_group_process_chunks(self._h, n, ids.ctypes.data, ptrs, lens, status.ctypes.data)
        if rc and raise_on_error:
            raise AmiraError(rc, self._err())
        self.last_status = status[:n].tolist()
        return [self.transcript(int(s)) for s in ids]

    # -- incremental mode (amira_b200.h: amira_stream_group_set_incremental): each chunk costs its own frames
    def set_incremental(self, enable: bool = True):
        rc = self._L.amira_stream_group_set_incremental(self._h, int(enable))
        if rc:
            raise AmiraError(rc, self._err())

    def flush(self, streams, raise_on_error: bool = True):
        """End of the listed streams: the remaining frames (reflect padding of the true end); returns their transcripts."""
        ids = np.ascontiguousarray(streams, dtype=np.int32)
        status = np.zeros(max(ids.size, 1), np.int32)
        rc = self._L.amira_stream_group_flush(self._h, int(ids.size), ids.ctypes.data, status.ctypes.data)
        if rc and raise_on_error:
            raise AmiraError(rc, self._err())
        self.last_status = status[:ids.size].tolist()
        return [self.transcript(int(s)) for s in ids]

    def progress(self, stream: int):
        """(samples received, log-mel frames emitted, encoder frames decoded) of one stream."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _check(self._L.amira_stream_group_progress(self._h, stream, C.byref(a), C.byref(b), C.byref(c)), "progress")
        return int(a.value), int(b.value), int(c.value)

    def transcript(self, stream: int) -> str:
        n = C.c_int32(0)
        _check(self._L.amira_stream_group_transcript(self._h, stream, None, 0, C.byref(n)), "transcript")
        out = C.create_string_buffer(n.value + 1)
        _check(self._L.amira_stream_group_transcript(self._h, stream, out, n.value + 1, C.byref(n)), "transcript")
        return out.raw[:n.value].decode("utf-8")

    def tokens(self, stream: int) -> list:
        n = C.c_int32(0)
        _check(self._L.amira_stream_group_tokens(self._h, stream, None, 0, C.byref(n)), "tokens")
        t = np.zeros(max(n.value, 1), np.int32)
        _check(self._L.amira_stream_group_tokens(self._h, stream, t.ctypes.data, n.value, C.byref(n)), "tokens")
        return t[:n.value].tolist()

    def audio_length(self, stream: int) -> float:
        s = C.c_float(0)
        _check(self._L.amira_stream_group_audio_length(self._h, stream, C.byref(s)), "audio_length")
        return float(s.value)

    def clear(self, stream: int):
        _check(self._L.amira_stream_group_clear(self._h, stream), "clear")

    def process_batch(self, stream: int, audio_bytes: bytes) -> Transcription:
        """IncrementalAsr::process_batch (src/asr/incremental.rs:267-292)"""
        t = _Transcription()
        toks = np.zeros(1 << 16, np.int32)
        text = C.create_string_buffer(1 << 16)
        b = np.frombuffer(bytes(audio_bytes), dtype=np.uint8)
        rc = self._L.amira_stream_group_process_batch(self._h, stream, b.ctypes.data if b.size else None, b.size, C.byref(t),
                                                      toks.ctypes.data, toks.size, text, len(text))
        if rc:
            raise AmiraError(rc, self._err())
        return Transcription(text.value.decode("utf-8", "replace"), toks[:t.n_tokens].tolist(), t.audio_length_samples,
                             t.features_length, t.encoded_length)

    def stats(self):
        """(process_stream_samples calls served, batched rounds run)"""
        a, b = C.c_int64(0), C.c_int64(0)
        self._L.amira_stream_group_stats(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.amira_stream_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IncrementalAsr:
    """One stream: the reference's object, same method names (src/asr/incremental.rs:35-298)."""

    def __init__(self, pipeline: B200AsrPipeline, chunk_size: float = CHUNK_SIZE, leading_context: float = LEADING_CONTEXT,
                 trailing_context: float = TRAILING_CONTEXT, buffer_capacity: float = BUFFER_CAPACITY_SECONDS):
        self._g = StreamGroup(pipeline, 1, chunk_size, leading_context, trailing_context, buffer_capacity)

    def process_chunk(self, audio_bytes: bytes) -> str:
        return self._g.process_chunks([0], [audio_bytes])[0]

    def process_batch(self, audio_bytes: bytes) -> Transcription:
        return self._g.process_batch(0, audio_bytes)

    def clear(self):
        self._g.clear(0)

    def audio_length(self) -> float:
        return self._g.audio_length(0)

    @property
    def tokens(self) -> list:
        return self._g.tokens(0)

    def close(self):
        self._g.close()
