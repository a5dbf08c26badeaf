"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's streaming orchestrator — SURVEY.md 8(f1).

Only tests/ may import this file; the product path (csrc/host_stream.cpp) never does.

Restates, function by function:
  * window_sequence / OverlappingAudioBuffer        src/asr/audio.rs:72-132, 134-293
  * mean_amplitude_optimized                        src/performance_opts.rs:35-60
  * levenshtein_distance .. weave_transcript_segs   src/asr/weaving.rs:16-280
  * is_overlap_silence                              src/asr/weaving.rs:285-313
  * IncrementalAsr                                  src/asr/incremental.rs:21-298
  * bytes_to_f32_samples                            src/asr/audio.rs:18-26

Rust semantics kept on purpose: `str::len()` is BYTES while `chars()` are Unicode scalars; all float arithmetic is f32
(numpy.float32 here); usize arithmetic wraps (release build) — the short-last-window branch of window_sequence relies on
two wraps cancelling (audio.rs:112-115).  PARITY UNPINNED: the reference ships no tests for these functions; this file
and the C++ implementation are two independent restatements of the cited lines and are checked against each other.
"""
from __future__ import annotations

import functools
import math

import numpy as np

F = np.float32
EXPECTED_SILENCE_RATIO = F(2.0)   # src/asr/types.rs:16
MAX_ALIGN_DIST = F(0.6)           # src/asr/types.rs:18
ALPHA = F(0.1)                    # src/asr/types.rs:20
W2V_SAMPLE_RATE = 16000           # src/asr/types.rs:22
MIN_ALIGNMENT_SCORE = F(0.01)     # src/asr/incremental.rs:19
U64 = (1 << 64) - 1


def _blen(s: str) -> int:
    return len(s.encode("utf-8"))


def _as_usize(x) -> int:
    """Rust `f32 as usize`: truncation toward zero, saturating, NaN -> 0."""
    x = float(x)
    if math.isnan(x) or x <= 0.0:
        return 0
    if x >= 18446744073709551616.0:
        return U64
    return int(x)


# ------------------------------------------------------------------------------------------------ audio.rs
def bytes_to_f32_samples(audio_bytes: bytes) -> np.ndarray:
    """src/asr/audio.rs:18-26 — chunks_exact(2): an odd trailing byte is dropped."""
    n = len(audio_bytes) // 2
    return (np.frombuffer(audio_bytes[:2 * n], dtype="<i2").astype(np.float32) / F(32768.0)).astype(np.float32)


def mean_amplitude(samples: np.ndarray) -> np.float32:
    """src/performance_opts.rs:35-60 — one f32 accumulator, samples added in index order."""
    if len(samples) == 0:
        return F(0.0)
    acc = F(0.0)
    for v in np.abs(np.asarray(samples, dtype=np.float32)):
        acc = F(acc + v)
    return F(acc / F(len(samples)))


def window_sequence(total_len: int, window_size: int, leading: int, trailing: int):
    """src/asr/audio.rs:98-132 — yields (source (start, end), target (start, end), overlap_ratio)."""
    out = []
    consumed = 0
    while consumed < total_len:
        start = consumed
        end = min(total_len, consumed + window_size)
        offset = min(leading, consumed)
        overlap = trailing + leading
        if end < total_len:
            consumed = (end - leading - trailing) & U64
        else:
            consumed = end
            if end - start < window_size:
                new_start = max(0, (end - window_size) & U64)  # usize: wraps when end < window_size
                overlap = (overlap + ((start - new_start) & U64)) & U64
        out.append(((start, end), ((start + offset) & U64, end), F(F(overlap) / F(window_size))))
    return out


class OverlappingAudioBuffer:
    """src/asr/audio.rs:134-293"""

    def __init__(self, capacity: int, chunk_size: float, leading_context: float, trailing_context: float):
        self.buffer = np.zeros(capacity, np.float32)
        self.length = 0
        self.capacity = capacity
        self.chunk_size = _as_usize(F(F(chunk_size) * F(W2V_SAMPLE_RATE)))
        self.leading_context = _as_usize(F(F(leading_context) * F(W2V_SAMPLE_RATE)))
        self.trailing_context = _as_usize(F(F(trailing_context) * F(W2V_SAMPLE_RATE)))
        self.mean_amplitude = F(0.0)

    def add_samples(self, samples: np.ndarray):
        n = len(samples)
        if self.length + n > self.capacity:
            keep = min(self.leading_context, self.length)
            start = self.length - keep
            if keep > 0:
                self.buffer[:keep] = self.buffer[start:self.length].copy()
            self.length = keep
        s, e = self.length, self.length + n
        if e <= self.capacity:
            self.buffer[s:e] = samples
            self.length = e
            new_amp = mean_amplitude(samples)
            if self.mean_amplitude == F(0.0):
                self.mean_amplitude = new_amp
            else:
                self.mean_amplitude = F(F(F(0.7) * self.mean_amplitude) + F(F(0.3) * new_amp))
        else:  # truncation: the mean amplitude is NOT updated (audio.rs:236-241)
            avail = self.capacity - s
            self.buffer[s:self.capacity] = samples[:avail]
            self.length = self.capacity

    def get_window(self) -> np.ndarray:
        return self.buffer[:self.length]

    def overlapping_windows(self):
        return window_sequence(self.length, self.chunk_size + self.leading_context + self.trailing_context,
                               self.leading_context, self.trailing_context)

    def is_empty(self) -> bool:
        return self.length == 0

    def clear(self):
        self.length = 0
        self.mean_amplitude = F(0.0)

    def get_slice(self, sl) -> np.ndarray:
        return self.buffer[sl[0]:min(sl[1], self.length)]


# ------------------------------------------------------------------------------------------------ weaving.rs
@functools.lru_cache(maxsize=1 << 18)
def levenshtein_distance(s1: str, s2: str) -> int:
    """src/asr/weaving.rs:16-65 — note the empty-string shortcuts return BYTE lengths.  (The DP runs one numpy row at a time
    and results are memoised: weave_transcript_segs calls this O(overlap^2) times on heavily repeated pairs.)"""
    if s1 == s2:
        return 0
    if not s1:
        return _blen(s2)
    if not s2:
        return _blen(s1)
    if len(s1) > len(s2):  # the distance is symmetric: loop over the shorter string, vectorise over the longer
        s1, s2 = s2, s1
    b = np.frombuffer(s2.encode("utf-32-le"), dtype=np.uint32)
    prev = np.arange(len(s2) + 1, dtype=np.int64)
    idx = np.arange(len(s2) + 1, dtype=np.int64)
    for i, ch in enumerate(s1, start=1):
        cost = (b != ord(ch)).astype(np.int64)
        cand = np.empty(len(s2) + 1, dtype=np.int64)
        cand[0] = i
        cand[1:] = np.minimum(prev[1:] + 1, prev[:-1] + cost)      # deletion / substitution
        # insertion: cur[j] = min_k<=j (cand[k] + (j - k)) — a running minimum of cand[k] - k
        prev = np.minimum.accumulate(cand - idx) + idx
    return int(prev[len(s2)])


def word_distance(first: str, second: str) -> np.float32:
    """src/asr/weaving.rs:71-86"""
    if first == second:
        return F(0.0)
    fl, sl = _blen(first), _blen(second)
    if fl == 0 and sl == 0:
        return F(0.0)
    return F(F(F(2.0) * F(levenshtein_distance(first, second))) / F(fl + sl))


def overlap_prior(first: str, second: str, overlap: int, percent_time) -> np.float32:
    """src/asr/weaving.rs:92-104"""
    with np.errstate(all="ignore"):
        fl, sl = F(_blen(first)), F(_blen(second))
        mu = F(F(F(F(fl * F(3.0)) + F(sl * F(2.0))) * F(percent_time)) / F(5.0))
        sigma = F(mu / F(2.0))
        diff = F(F(F(overlap) - mu) / sigma)
        exponent = F(F(F(-0.5) * diff) * diff)
        norm = F(sigma * np.sqrt(F(F(2.0) * F(math.pi))))
        return F(np.exp(exponent) / norm)


def dist_score(dist) -> np.float32:
    """src/asr/weaving.rs:109-111"""
    return F(F(F(1.0) / F(F(dist) + ALPHA)) - F(F(1.0) / F(F(1.0) + ALPHA)))


def _first_end(first: str, overlap: int) -> str:
    """`first.char_indices().nth_back(count.saturating_sub(overlap))` -> `&first[idx..]` (weaving.rs:122-128, 154-160):
    the suffix starting at char min(overlap, count) - 1 — literal, including what looks like an off-by-intent."""
    n = len(first)
    k = max(n - overlap, 0)
    if k >= n:
        return first
    return first[n - 1 - k:]


def _second_start(second: str, overlap: int) -> str:
    """`second.char_indices().nth(overlap.saturating_sub(1))` -> `&second[..idx]`, whole string if out of range"""
    k = max(overlap - 1, 0)
    return second[:k] if k < len(second) else second


def align_score(first: str, second: str, overlap: int, percent_time) -> np.float32:
    """src/asr/weaving.rs:117-142"""
    if _blen(first) < overlap or _blen(second) < overlap:
        return F(0.0)
    dist = word_distance(_first_end(first, overlap), _second_start(second, overlap))
    if dist > MAX_ALIGN_DIST:
        return F(0.0)
    with np.errstate(all="ignore"):
        return F(overlap_prior(first, second, overlap, percent_time) * dist_score(dist))


def trim_align_score(first: str, second: str, overlap: int) -> np.float32:
    """src/asr/weaving.rs:148-174"""
    if not first or not second or overlap == 0:
        return F(0.0)
    dist = word_distance(_first_end(first, overlap), _second_start(second, overlap))
    if dist > MAX_ALIGN_DIST:
        return F(0.0)
    return F(F(F(1.0) - dist) * np.sqrt(F(overlap)))


def best_alignment(first: str, second: str, percent_time):
    """src/asr/weaving.rs:180-203"""
    best_score, best_overlap = F(0.0), 0
    fl, sl = len(first), len(second)
    if fl == 0 or sl == 0:
        return 0, F(0.0)
    max_overlap = min(fl, _as_usize(F(F(sl) * F(1.25))))
    for overlap in range(1, max_overlap + 1):
        score = align_score(first, second, overlap, percent_time)
        if score > best_score:
            best_score, best_overlap = score, overlap
    return best_overlap, best_score


def weave_transcript_segs(first_seg: str, second_seg: str, percent_time_overlap, min_alignment_score=MIN_ALIGNMENT_SCORE) -> str:
    """src/asr/weaving.rs:209-280"""
    overlap, a_score = best_alignment(first_seg, second_seg, percent_time_overlap)
    if overlap == 0 or a_score < F(min_alignment_score):
        return first_seg + " " + second_seg
    best_score, best_trim = F(0.0), (0, 0)
    n1, n2 = len(first_seg), len(second_seg)
    for idx in range(overlap + 1):
        left_start = 0 if idx >= overlap else max(n1 - (overlap - idx), 0)
        left = first_seg[left_start:] if left_start < n1 else first_seg  # nth() None -> byte index 0
        for idx2 in range(overlap + 1):
            right = second_seg[:min(overlap, n2)]
            adjusted = max(overlap * 2 - (idx + idx2), 0)
            score = trim_align_score(left, right, adjusted)
            if score > best_score:
                best_score, best_trim = score, (idx, idx2)
    if best_trim[0] >= overlap:
        first_keep = n1
    else:
        first_keep = min(max(n1 - (overlap - best_trim[0]), 0), n1)
    second_trim = best_trim[1] if best_trim[1] < n2 else 0  # nth() None -> 0
    return first_seg[:first_keep] + second_seg[second_trim:]


def is_overlap_silence(overlap_audio: np.ndarray, mean_amp) -> bool:
    """src/asr/weaving.rs:285-313 — per-window f32 sums, added in index order from 0.0"""
    a = np.asarray(overlap_audio, dtype=np.float32)
    if a.size == 0:
        return True
    sq = (a * a).astype(np.float32)
    w = min(800, sq.size)
    nw = sq.size - w + 1
    view = np.lib.stride_tricks.sliding_window_view(sq, w)  # [nw, w]
    acc = np.zeros(nw, np.float32)
    for k in range(w):
        acc = (acc + view[:, k]).astype(np.float32)
    avg = (acc / F(w)).astype(np.float32)
    max_energy = F(0.0)
    for v in avg:  # f32::max ignores NaN
        if v > max_energy:
            max_energy = F(v)
    peak = F(np.sqrt(max_energy))
    return bool(peak < F(F(mean_amp) / EXPECTED_SILENCE_RATIO))


# ------------------------------------------------------------------------------------------------ incremental.rs
def sample_index_to_logit_index(idx: int) -> int:
    """src/asr/incremental.rs:27-29"""
    return _as_usize(F(F(F(idx) * F(299.0)) / F(96000.0)))


class IncrementalAsr:
    """src/asr/incremental.rs:35-298.  `pipeline` is any object with
    process_stream_samples(samples f32, state) -> (text, tokens) and process_batch(bytes) -> (text, tokens, n_samples, flen, elen);
    `new_state()` makes a zero DecoderState."""

    def __init__(self, pipeline, new_state, chunk_size=2.0, leading_context=1.0, trailing_context=0.5, buffer_capacity=10.0):
        self.pipeline = pipeline
        self.new_state = new_state
        capacity = _as_usize(F(F(buffer_capacity) * F(W2V_SAMPLE_RATE)))
        self.audio_buffer = OverlappingAudioBuffer(capacity, chunk_size, leading_context, trailing_context)
        self.token_ids: list[int] = []
        self.transcript = ""
        self.mean_amplitude = F(0.0)
        self.decoder_state = new_state()
        self.chunk_size = F(chunk_size)
        self.n_pipeline_calls = 0

    def clear(self):
        self.audio_buffer.clear()
        self.token_ids = []
        self.transcript = ""
        self.mean_amplitude = F(0.0)
        self.decoder_state = self.new_state()

    def process_chunk(self, audio_bytes: bytes) -> str:
        """:111-129"""
        self.audio_buffer.add_samples(bytes_to_f32_samples(audio_bytes))
        self.mean_amplitude = self.audio_buffer.mean_amplitude
        if not self.audio_buffer.is_empty():
            self._process_buffered_audio()
        return self.transcript

    def _process_buffered_audio(self):
        """:135-170"""
        if not self.token_ids:
            text, tokens = self.pipeline.process_stream_samples(self.audio_buffer.get_window().copy(), self.decoder_state)
            self.n_pipeline_calls += 1
            self.token_ids = list(tokens)
            self.transcript = text
            return
        for src, tgt, overlap in self.audio_buffer.overlapping_windows():
            chunk = self.audio_buffer.get_slice(src).copy()
            text, tokens = self.pipeline.process_stream_samples(chunk, self.decoder_state)
            self.n_pipeline_calls += 1
            self._accumulate(text, list(tokens), tgt, overlap)

    def _accumulate(self, segment: str, tokens: list, target_slice, overlap):
        """:181-258"""
        if self.transcript == "":
            self.transcript = segment
            self.token_ids = list(tokens)
            return
        chunk = _as_usize(F(F(F(overlap) * self.chunk_size) * F(W2V_SAMPLE_RATE)))
        if chunk > 0:
            win = self.audio_buffer.get_window()
            start = max(len(win) - chunk, 0)
            silence = is_overlap_silence(win[start:], self.mean_amplitude)
        else:
            silence = False
        if silence:
            self.transcript = self.transcript + " " + segment
        else:
            self.transcript = weave_transcript_segs(self.transcript, segment, overlap, MIN_ALIGNMENT_SCORE)
        ls, le = sample_index_to_logit_index(target_slice[0]), sample_index_to_logit_index(target_slice[1])
        if len(self.token_ids) < le:
            self.token_ids.extend([0] * (le - len(self.token_ids)))
        n_copy = min(len(tokens), (le - ls) & U64)
        if n_copy > 0 and ls < len(self.token_ids):
            end = min(ls + n_copy, len(self.token_ids))
            assert end - ls == n_copy  # copy_from_slice would panic otherwise; unreachable after the resize above
            self.token_ids[ls:end] = tokens[:n_copy]

    def process_batch(self, audio_bytes: bytes):
        """:267-292 -> (text, tokens, audio_length_samples, features_length, encoded_length)"""
        self.clear()
        samples = bytes_to_f32_samples(audio_bytes)
        if F(F(len(samples)) / F(W2V_SAMPLE_RATE)) <= self.chunk_size:
            return self.pipeline.process_batch(audio_bytes)
        self.audio_buffer.add_samples(samples)
        self._process_buffered_audio()
        return self.transcript, list(self.token_ids), len(samples), 0, 0

    def audio_length(self) -> np.float32:
        return F(F(self.audio_buffer.length) / F(W2V_SAMPLE_RATE))
