// host_stream.cpp — the streaming orchestrator above the pipeline: C++ mirror of the reference's
// IncrementalAsr (src/asr/incremental.rs:35-298), OverlappingAudioBuffer / window_sequence (src/asr/audio.rs:72-293),
// transcript weaving (src/asr/weaving.rs:16-280) and overlap-silence detection (src/asr/weaving.rs:285-313).
//
// The reference runs one IncrementalAsr per WebSocket and one Triton round trip per window per decode step.  Here a
// *stream group* holds many sessions and `amira_stream_group_process_chunks` advances all of them by one chunk: window k of
// every session goes through ONE batched front-end launch, the injected encoder, and ONE persistent decode launch (each
// session's LSTM state carried from its window k-1), then each session weaves its transcript on the host.  Sessions are
// independent, so the result per session equals the reference's one-at-a-time processing.
//
// Rust semantics kept literally: str::len() is bytes while chars() are Unicode scalars; all float arithmetic is f32;
// usize arithmetic wraps as in a release build (window_sequence's short-last-window branch, audio.rs:112-115).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <condition_variable>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include "host_common.h"

namespace {

using u32s = std::u32string;

// ---- UTF-8 <-> scalar values (Rust strings are valid UTF-8; malformed bytes decode as one char each) ----
u32s decode_utf8(const char *s) {
    u32s out;
    const unsigned char *p = reinterpret_cast<const unsigned char *>(s ? s : "");
    while (*p) {
        uint32_t c = *p;
        int extra = c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : c >= 0xC0 ? 1 : 0;
        bool ok = c < 0x80 || (c >= 0xC2 && c <= 0xF4);
        for (int k = 1; k <= extra && ok; ++k) ok = (p[k] & 0xC0) == 0x80;
        if (!ok || extra == 0) { out.push_back(c); ++p; continue; }
        uint32_t v = c & (0x3F >> extra);
        for (int k = 1; k <= extra; ++k) v = (v << 6) | (p[k] & 0x3F);
        out.push_back(v);
        p += extra + 1;
    }
    return out;
}
size_t cp_bytes(char32_t c) { return c < 0x80 ? 1 : c < 0x800 ? 2 : c < 0x10000 ? 3 : 4; }
size_t blen(const char32_t *s, size_t n) {  // str::len(): bytes
    size_t b = 0;
    for (size_t i = 0; i < n; ++i) b += cp_bytes(s[i]);
    return b;
}
std::string encode_utf8(const u32s &s) {
    std::string out;
    for (char32_t c : s) {
        if (c < 0x80) out += (char)c;
        else if (c < 0x800) { out += (char)(0xC0 | (c >> 6)); out += (char)(0x80 | (c & 0x3F)); }
        else if (c < 0x10000) { out += (char)(0xE0 | (c >> 12)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); }
        else { out += (char)(0xF0 | (c >> 18)); out += (char)(0x80 | ((c >> 12) & 0x3F)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); }
    }
    return out;
}
struct View {  // a &str: chars [p, p + n)
    const char32_t *p;
    size_t n;
    size_t bytes() const { return blen(p, n); }
    bool eq(const View &o) const { return n == o.n && std::equal(p, p + n, o.p); }
};
size_t f32_as_usize(float x) {  // Rust `as usize`: truncating, saturating, NaN -> 0
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)x;
}

// ---- weaving.rs ----
constexpr float kExpectedSilenceRatio = 2.0f, kMaxAlignDist = 0.6f, kAlpha = 0.1f;  // src/asr/types.rs:16-20

size_t levenshtein_distance(View a, View b) {  // weaving.rs:16-65 (empty-string shortcuts return BYTE lengths)
    if (a.eq(b)) return 0;
    if (a.n == 0) return b.bytes();
    if (b.n == 0) return a.bytes();
    std::vector<size_t> prev(b.n + 1), cur(b.n + 1);
    for (size_t j = 0; j <= b.n; ++j) prev[j] = j;
    for (size_t i = 1; i <= a.n; ++i) {
        cur[0] = i;
        for (size_t j = 1; j <= b.n; ++j) {
            const size_t cost = a.p[i - 1] == b.p[j - 1] ? 0 : 1;
            cur[j] = std::min(prev[j] + 1, std::min(cur[j - 1] + 1, prev[j - 1] + cost));
        }
        std::swap(prev, cur);
    }
    return prev[b.n];
}
float word_distance(View a, View b) {  // weaving.rs:71-86
    if (a.eq(b)) return 0.0f;
    const size_t al = a.bytes(), bl = b.bytes();
    if (al == 0 && bl == 0) return 0.0f;
    const float d = (float)levenshtein_distance(a, b);
    return 2.0f * d / (float)(al + bl);
}
float overlap_prior(View first, View second, size_t overlap, float percent_time) {  // weaving.rs:92-104
    const float mu = ((float)first.bytes() * 3.0f + (float)second.bytes() * 2.0f) * percent_time / 5.0f;
    const float sigma = mu / 2.0f;
    const float diff = ((float)overlap - mu) / sigma;
    const float exponent = -0.5f * diff * diff;
    const float normalization = sigma * std::sqrt(2.0f * 3.14159265358979323846f);
    return std::exp(exponent) / normalization;
}
float dist_score(float dist) { return 1.0f / (dist + kAlpha) - 1.0f / (1.0f + kAlpha); }  // weaving.rs:109-111
// `first.char_indices().nth_back(count.saturating_sub(overlap))` -> &first[idx..] (weaving.rs:122-128, 154-160)
View first_end(View first, size_t overlap) {
    const size_t n = first.n, k = n > overlap ? n - overlap : 0;
    if (k >= n) return first;
    const size_t i = n - 1 - k;
    return {first.p + i, n - i};
}
// `second.char_indices().nth(overlap.saturating_sub(1))` -> &second[..idx]; out of range -> whole string
View second_start(View second, size_t overlap) {
    const size_t k = overlap > 0 ? overlap - 1 : 0;
    return k < second.n ? View{second.p, k} : second;
}
float align_score(View first, View second, size_t overlap, float percent_time) {  // weaving.rs:117-142
    if (first.bytes() < overlap || second.bytes() < overlap) return 0.0f;
    const float dist = word_distance(first_end(first, overlap), second_start(second, overlap));
    if (dist > kMaxAlignDist) return 0.0f;
    return overlap_prior(first, second, overlap, percent_time) * dist_score(dist);
}
float trim_align_score(View first, View second, size_t overlap) {  // weaving.rs:148-174
    if (first.n == 0 || second.n == 0 || overlap == 0) return 0.0f;
    const float dist = word_distance(first_end(first, overlap), second_start(second, overlap));
    if (dist > kMaxAlignDist) return 0.0f;
    return (1.0f - dist) * std::sqrt((float)overlap);
}
void best_alignment(View first, View second, float percent_time, size_t *best_overlap, float *best_score) {  // weaving.rs:180-203
    *best_overlap = 0;
    *best_score = 0.0f;
    if (first.n == 0 || second.n == 0) return;
    const size_t max_overlap = std::min(first.n, f32_as_usize((float)second.n * 1.25f));
    for (size_t overlap = 1; overlap <= max_overlap; ++overlap) {
        const float score = align_score(first, second, overlap, percent_time);
        if (score > *best_score) { *best_score = score; *best_overlap = overlap; }
    }
}
u32s weave_transcript_segs(const u32s &a, const u32s &b, float percent_time_overlap, float min_alignment_score) {  // weaving.rs:209-280
    const View first{a.data(), a.size()}, second{b.data(), b.size()};
    size_t overlap;
    float a_score;
    best_alignment(first, second, percent_time_overlap, &overlap, &a_score);
    if (overlap == 0 || a_score < min_alignment_score) return a + U" " + b;
    float best_score = 0.0f;
    size_t trim0 = 0, trim1 = 0;
    const size_t n1 = a.size(), n2 = b.size();
    for (size_t idx = 0; idx <= overlap; ++idx) {
        size_t left_start = idx >= overlap ? 0 : (n1 > overlap - idx ? n1 - (overlap - idx) : 0);
        if (left_start >= n1) left_start = 0;  // nth() == None -> byte index 0
        const View left{first.p + left_start, n1 - left_start};
        for (size_t idx2 = 0; idx2 <= overlap; ++idx2) {
            const View right{second.p, std::min(overlap, n2)};
            const size_t adjusted = overlap * 2 > idx + idx2 ? overlap * 2 - (idx + idx2) : 0;
            const float score = trim_align_score(left, right, adjusted);
            if (score > best_score) { best_score = score; trim0 = idx; trim1 = idx2; }
        }
    }
    size_t first_keep;
    if (trim0 >= overlap) first_keep = n1;
    else first_keep = std::min(n1 > overlap - trim0 ? n1 - (overlap - trim0) : 0, n1);
    const size_t second_trim = trim1 < n2 ? trim1 : 0;
    return a.substr(0, first_keep) + b.substr(second_trim);
}
bool is_overlap_silence(const float *audio, size_t n, float mean_amplitude) {  // weaving.rs:285-313
    if (n == 0) return true;
    const size_t w = std::min<size_t>(800, n);
    float max_energy = 0.0f;
    for (size_t i = 0; i + w <= n; ++i) {
        float sum = 0.0f;  // per-window sums in index order, as `iter().sum()` adds them
        for (size_t k = 0; k < w; ++k) sum += audio[i + k] * audio[i + k];
        const float avg = sum / (float)w;
        if (avg > max_energy) max_energy = avg;  // f32::max ignores NaN
    }
    return std::sqrt(max_energy) < mean_amplitude / kExpectedSilenceRatio;
}

// ---- audio.rs ----
float mean_amplitude_of(const float *s, size_t n) {  // performance_opts.rs:35-60: one accumulator, index order
    if (n == 0) return 0.0f;
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += std::fabs(s[i]);
    return sum / (float)n;
}
struct Window { size_t src_start, src_end, tgt_start, tgt_end; float overlap; };
std::vector<Window> window_sequence(size_t total_len, size_t window_size, size_t leading, size_t trailing) {  // audio.rs:98-132
    std::vector<Window> out;
    size_t consumed = 0;
    while (consumed < total_len) {
        const size_t start = consumed, end = std::min(total_len, consumed + window_size);
        const size_t offset = std::min(leading, consumed);
        size_t overlap = trailing + leading;
        if (end < total_len) {
            consumed = end - leading - trailing;
        } else {
            consumed = end;
            if (end - start < window_size) {
                const size_t new_start = end - window_size;  // std::cmp::max(0, ..) on usize: wraps when end < window_size,
                overlap += start - new_start;                // and wraps back here (release-build arithmetic)
            }
        }
        out.push_back({start, end, start + offset, end, (float)overlap / (float)window_size});
        if (out.size() > (1u << 20)) break;  // degenerate parameters (window <= contexts) would never advance
    }
    return out;
}
struct OverlappingAudioBuffer {  // audio.rs:134-293
    std::vector<float> buffer;
    size_t length = 0, capacity = 0, leading = 0, trailing = 0, chunk = 0;
    float mean_amplitude = 0.0f;
    void init(size_t cap, float chunk_size, float leading_context, float trailing_context) {
        buffer.assign(cap, 0.0f);
        capacity = cap;
        chunk = f32_as_usize(chunk_size * 16000.0f);
        leading = f32_as_usize(leading_context * 16000.0f);
        trailing = f32_as_usize(trailing_context * 16000.0f);
    }
    void add_samples(const float *s, size_t n) {
        if (length + n > capacity) {
            const size_t keep = std::min(leading, length), start = length - keep;
            if (keep > 0) std::memmove(buffer.data(), buffer.data() + start, sizeof(float) * keep);
            length = keep;
        }
        const size_t s0 = length, e0 = s0 + n;
        if (e0 <= capacity) {
            if (n) std::memcpy(buffer.data() + s0, s, sizeof(float) * n);
            length = e0;
            const float amp = mean_amplitude_of(s, n);
            mean_amplitude = mean_amplitude == 0.0f ? amp : 0.7f * mean_amplitude + 0.3f * amp;
        } else {  // truncation; the mean amplitude is not updated (audio.rs:236-241)
            const size_t avail = capacity - s0;
            if (avail) std::memcpy(buffer.data() + s0, s, sizeof(float) * avail);
            length = capacity;
        }
    }
    std::vector<Window> overlapping_windows() const { return window_sequence(length, chunk + leading + trailing, leading, trailing); }
    void clear() { length = 0; mean_amplitude = 0.0f; }
};

size_t sample_index_to_logit_index(size_t idx) { return f32_as_usize(((float)idx * 299.0f) / 96000.0f); }  // incremental.rs:27-29
constexpr float kMinAlignmentScore = 0.01f;  // incremental.rs:19

// ---- one IncrementalAsr (incremental.rs:35-61) ----
struct Session {
    OverlappingAudioBuffer audio;
    std::vector<int32_t> token_ids;   // AccumulatedPredictions (types.rs:185-212)
    u32s transcript;
    float acc_mean_amplitude = 0.0f;
    std::vector<float> s1, s2;        // DecoderState [2][1][640] (types.rs:159-183)
    float chunk_size = 2.0f;
    // ---- incremental mode: what a stream carries between chunks ----
    std::vector<int16_t> hist;        // audio tail: samples [hist_start, n_total) of the stream
    int64_t hist_start = 0, n_total = 0;
    int64_t t_next = 0;               // first log-mel frame not emitted yet
    int64_t enc_frames = 0;
    bool flushed = false;
    std::vector<double> st_mean = std::vector<double>(AMIRA_N_MELS, 0.0), st_m2 = std::vector<double>(AMIRA_N_MELS, 0.0);  // running
                                      // per-feature statistics over the st_count frames emitted so far
    int64_t st_count = 0;
    int32_t last_token = AMIRA_BLANK_ID;
    bool transcript_stale = false;    // incremental mode: `transcript` is rebuilt from token_ids when somebody asks for it
    void reset_state() { s1.assign(2 * AMIRA_STATE_SIZE, 0.0f); s2.assign(2 * AMIRA_STATE_SIZE, 0.0f); }
    void clear() {  // incremental.rs:99-103
        audio.clear();
        token_ids.clear();
        transcript.clear();
        acc_mean_amplitude = 0.0f;
        reset_state();
        hist.clear();
        hist_start = n_total = t_next = enc_frames = st_count = 0;
        flushed = false;
        st_mean.assign(AMIRA_N_MELS, 0.0);
        st_m2.assign(AMIRA_N_MELS, 0.0);
        last_token = AMIRA_BLANK_ID;
        transcript_stale = false;
    }
};

struct Job {  // one process_stream_samples call of one session
    Session *s;
    size_t src_start, src_end, tgt_start, tgt_end;
    float overlap;
    bool first;  // the "no tokens yet" branch (incremental.rs:139-150): whole window, result replaces the accumulated state
    std::vector<int32_t> tokens;
    std::string text;
    int64_t flen = 0, elen = 0;
    int32_t rc = AMIRA_OK;
};

}  // namespace

// Host scratch of the incremental rounds in page-locked memory: the library copies pinned host buffers asynchronously at the
// link's rate, pageable ones through the driver's bounce buffer (measured on a 1024-stream tick: front end 1.5 -> 0.8 ms, resumed
// decode 3.9 -> 2.6 ms).  Falls back to pageable memory where cudaMallocHost fails.  Contents survive growth.
template <class T>
struct PinVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    PinVec() = default;
    PinVec(const PinVec &) = delete;
    PinVec &operator=(const PinVec &) = delete;
    ~PinVec() { release(p, pinned); }
    static void release(T *q, bool pin) {
        if (!q) return;
        if (pin) { cudaFreeHost(q); cudaGetLastError(); }
        else std::free(q);
    }
    void resize(size_t m) {
        if (m > cap) {
            const size_t want = m + m / 4 + 64;
            T *q = nullptr;
            bool pin = cudaMallocHost(reinterpret_cast<void **>(&q), want * sizeof(T)) == cudaSuccess;
            if (!pin) {
                cudaGetLastError();
                q = static_cast<T *>(std::malloc(want * sizeof(T)));
                if (!q) throw std::bad_alloc();
            }
            if (n) std::memcpy(q, p, n * sizeof(T));
            release(p, pinned);
            p = q; cap = want; pinned = pin;
        }
        n = m;
    }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
};

// device memory of the group (the decoder states of the resident streams): the decode entries use device pointers in place
struct DevVec {
    float *p = nullptr;
    size_t cap = 0;
    DevVec() = default;
    DevVec(const DevVec &) = delete;
    DevVec &operator=(const DevVec &) = delete;
    ~DevVec() { if (p) { cudaFree(p); cudaGetLastError(); } }
    bool reserve(size_t n) {  // false: no device memory (the caller falls back to the host buffers)
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(reinterpret_cast<void **>(&p), (n + n / 4) * sizeof(float)) != cudaSuccess) { cudaGetLastError(); p = nullptr; return false; }
        cap = n + n / 4;
        return true;
    }
};

// A small pool of host threads for the per-session work of a round (sessions are independent).  The workers live as long as the
// library and sleep between rounds: starting fifteen threads per call cost more than the work they did, and the churn slowed
// the CUDA calls that followed on the calling thread (measured on a 1024-stream tick).  run(n, min_per_thread, f) calls
// f(lo, hi) over [0, n) split in up to 1 + workers pieces and returns when all are done; f must not throw.  One round at a time.
class HostPool {
  public:
    static HostPool &get() {
        static HostPool pool;
        return pool;
    }
    template <class F>
    void run(int n, int min_per_thread, F f) {
        const int T = std::max(1, std::min(n / std::max(min_per_thread, 1), (int)workers_.size() + 1));
        if (T <= 1) { f(0, n); return; }
        std::unique_lock<std::mutex> round(round_mu_);  // callers of different groups take turns
        std::function<void(int, int)> fn = f;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; pieces_ = T; next_ = 1; pending_ = T - 1;
            ++epoch_;
        }
        cv_.notify_all();
        f(0, (int)((long long)n / T));
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    HostPool() {
        const unsigned hw = std::thread::hardware_concurrency();
        const int n = std::max(0, std::min((int)(hw ? hw : 1), 16) - 1);
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && next_ < pieces_); });
            if (stop_) return;
            while (fn_ && next_ < pieces_) {
                const int t = next_++;
                const int lo = (int)((long long)n_ * t / pieces_), hi = (int)((long long)n_ * (t + 1) / pieces_);
                const std::function<void(int, int)> *fn = fn_;
                lk.unlock();
                (*fn)(lo, hi);
                lk.lock();
                if (--pending_ == 0) done_.notify_all();
            }
            seen = epoch_;
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, round_mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    int n_ = 0, pieces_ = 0, next_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

template <class F>
void parallel_for(int n, int min_per_thread, F f) {
    HostPool::get().run(n, min_per_thread, f);
}

struct amira_stream_group {
    amira_pipeline *p = nullptr;
    std::vector<std::unique_ptr<Session>> sessions;
    std::mutex mu;
    std::string err;
    int64_t n_pipeline_calls = 0, n_rounds = 0;
    bool incremental = false;
    // incremental mode: page-locked scratch of a round — the segments back to back, their log-mel blocks, the new frames of every
    // stream normalised (what the encoder reads), the encoder outputs packed, decoder states [2][B][640] x 2, last tokens, results
    PinVec<int16_t> pcm;
    PinVec<float> ifeat, ichunks, ienc, is1, is2;
    PinVec<int32_t> last, itokens, intok;
    std::vector<int64_t> coff;
    // the streams whose decoder state lives in is1 / is2 (row i = resident[i]) instead of their Session: as long as consecutive
    // rounds list the same streams in the same order the state never moves (1024 streams: 2 x 10.5 MB of memcpy per tick otherwise)
    std::vector<Session *> resident;
    DevVec ds1, ds2;                // ... and, when device memory is to be had, on the DEVICE: a tick then moves no state over PCIe
    bool resident_on_device = false;
    // batch scratch
    std::vector<float> wave, features, enc, s1, s2;
    std::vector<int64_t> woff, foff, eoff, flens, elens;
    std::vector<int32_t> tokens, ntok;
};

namespace {

int32_t gfail(amira_stream_group *g, int32_t code, const std::string &m) {
    if (g) g->err = m;
    return code;
}

// One round: job k of every session that has one — process_stream_samples (src/asr/pipeline.rs:416-443 -> :269-380) for all
// of them at once.  Per-job failures land in job.rc; a failure of a batched call fails every job of the round.
void run_round(amira_stream_group *g, std::vector<Job *> &jobs) {
    amira_pipeline *p = g->p;
    std::lock_guard<std::mutex> plock(p->mu);
    const int B = (int)jobs.size();
    auto fail_all = [&](int32_t rc) { for (Job *j : jobs) j->rc = rc; g->err = amira_last_error(p->ctx); };
    g->n_rounds++;
    g->n_pipeline_calls += B;
    // preprocessor, ragged in and out
    g->woff.assign((size_t)B + 1, 0);
    g->foff.assign((size_t)B + 1, 0);
    for (int i = 0; i < B; ++i) {
        const int64_t n = (int64_t)(jobs[(size_t)i]->src_end - jobs[(size_t)i]->src_start);
        int64_t fl = 0;
        amira_features_len(n, &fl);
        g->woff[(size_t)i + 1] = g->woff[(size_t)i] + n;
        g->foff[(size_t)i + 1] = g->foff[(size_t)i] + (int64_t)AMIRA_N_MELS * fl;
    }
    g->wave.resize((size_t)std::max<int64_t>(g->woff[(size_t)B], 1));
    for (int i = 0; i < B; ++i) {
        const Job *j = jobs[(size_t)i];
        std::memcpy(g->wave.data() + g->woff[(size_t)i], j->s->audio.buffer.data() + j->src_start, sizeof(float) * (j->src_end - j->src_start));
    }
    g->features.resize((size_t)std::max<int64_t>(g->foff[(size_t)B], 1));
    g->flens.assign((size_t)B, 0);
    int32_t rc = amira_preprocess_f32_packed(p->ctx, g->wave.data(), g->woff.data(), B, g->features.data(), g->foff.data(), g->flens.data());
    if (rc) return fail_all(rc);
    // encoder (out of scope, injected): per request [1][128][features_len] -> [1][1024][encoded_len]
    g->eoff.assign((size_t)B + 1, 0);
    g->elens.assign((size_t)B, 0);
    g->enc.clear();
    for (int i = 0; i < B; ++i) {
        Job *j = jobs[(size_t)i];
        j->flen = g->flens[(size_t)i];
        const float *e = nullptr;
        int64_t el = 0;
        if (!p->encoder || p->encoder(p->encoder_user, g->features.data() + g->foff[(size_t)i], j->flen, &e, &el) != 0 || el < 0 || (el > 0 && !e)) {
            j->rc = p->encoder ? AMIRA_ERR_UNKNOWN : AMIRA_ERR_NOT_READY;
            el = 0;
        }
        if (el > 0) g->enc.insert(g->enc.end(), e, e + (size_t)AMIRA_ENC_DIM * (size_t)el);  // the callback's buffer lives until its next call
        g->elens[(size_t)i] = el;
        j->elen = el;
        g->eoff[(size_t)i + 1] = g->eoff[(size_t)i] + (int64_t)AMIRA_ENC_DIM * el;
    }
    // greedy decode with each session's carried state ([2][B][640] in and out)
    int32_t cap = AMIRA_MAX_TOTAL_TOKENS;
    amira_ctx_max_total_tokens(p->ctx, &cap);
    const size_t H = AMIRA_STATE_SIZE;
    g->s1.resize(2 * (size_t)B * H);
    g->s2.resize(2 * (size_t)B * H);
    for (int i = 0; i < B; ++i)
        for (int l = 0; l < 2; ++l) {
            std::memcpy(g->s1.data() + ((size_t)l * B + i) * H, jobs[(size_t)i]->s->s1.data() + (size_t)l * H, sizeof(float) * H);
            std::memcpy(g->s2.data() + ((size_t)l * B + i) * H, jobs[(size_t)i]->s->s2.data() + (size_t)l * H, sizeof(float) * H);
        }
    g->tokens.assign((size_t)B * (size_t)cap, 0);
    g->ntok.assign((size_t)B, 0);
    if (g->eoff[(size_t)B] > 0) {
        rc = amira_greedy_decode_packed(p->ctx, g->enc.data(), g->eoff.data(), B, g->elens.data(), g->s1.data(), g->s2.data(),
                                        g->tokens.data(), g->ntok.data(), nullptr);
        if (rc && rc != AMIRA_ERR_DECODE_STEP) return fail_all(rc);
    }
    for (int i = 0; i < B; ++i) {
        Job *j = jobs[(size_t)i];
        if (j->rc) continue;
        if (g->ntok[(size_t)i] < 0) { j->rc = AMIRA_ERR_DECODE_STEP; continue; }  // "Decode step failed" (decoder_optimized.rs:148-152)
        if (j->elen > 0)  // a request without encoder frames never reaches the decoder: its state is untouched
            for (int l = 0; l < 2; ++l) {
                std::memcpy(j->s->s1.data() + (size_t)l * H, g->s1.data() + ((size_t)l * B + i) * H, sizeof(float) * H);
                std::memcpy(j->s->s2.data() + (size_t)l * H, g->s2.data() + ((size_t)l * B + i) * H, sizeof(float) * H);
            }
        const int32_t *tk = g->tokens.data() + (size_t)i * (size_t)cap;
        j->tokens.assign(tk, tk + g->ntok[(size_t)i]);
        j->text = p->vocab.decode(tk, g->ntok[(size_t)i]);
    }
}

// IncrementalAsr::accumulate_transcription (incremental.rs:181-258)
void accumulate(Session &s, const Job &j) {
    const u32s segment = decode_utf8(j.text.c_str());
    if (s.transcript.empty()) {
        s.transcript = segment;
        s.token_ids = j.tokens;
        return;
    }
    const size_t chunk = f32_as_usize(j.overlap * s.chunk_size * 16000.0f);
    bool silence = false;
    if (chunk > 0) {
        const size_t len = s.audio.length, start = len > chunk ? len - chunk : 0;
        silence = is_overlap_silence(s.audio.buffer.data() + start, len - start, s.acc_mean_amplitude);
    }
    if (silence) {
        s.transcript += U' ';
        s.transcript += segment;
    } else {
        s.transcript = weave_transcript_segs(s.transcript, segment, j.overlap, kMinAlignmentScore);
    }
    const size_t ls = sample_index_to_logit_index(j.tgt_start), le = sample_index_to_logit_index(j.tgt_end);
    if (s.token_ids.size() < le) s.token_ids.resize(le, 0);
    const size_t n_copy = std::min(j.tokens.size(), le - ls);
    if (n_copy > 0 && ls < s.token_ids.size()) {
        const size_t end = std::min(ls + n_copy, s.token_ids.size());
        std::copy(j.tokens.begin(), j.tokens.begin() + (end - ls), s.token_ids.begin() + ls);
    }
}

// process_buffered_audio (incremental.rs:135-170) of several sessions, advanced in lock step: round k = window k of each
int32_t process_buffered(amira_stream_group *g, const std::vector<Session *> &active, std::vector<int32_t> &rcs) {
    std::vector<std::vector<Job>> jobs(active.size());
    size_t rounds = 0;
    for (size_t i = 0; i < active.size(); ++i) {
        Session &s = *active[i];
        if (s.token_ids.empty()) {
            Job j{};
            j.s = &s; j.src_start = 0; j.src_end = s.audio.length; j.tgt_start = 0; j.tgt_end = s.audio.length; j.overlap = 0.f; j.first = true;
            jobs[i].push_back(std::move(j));
        } else {
            for (const Window &w : s.audio.overlapping_windows()) {
                Job j{};
                j.s = &s; j.src_start = w.src_start; j.src_end = std::min(w.src_end, s.audio.length);
                j.tgt_start = w.tgt_start; j.tgt_end = w.tgt_end; j.overlap = w.overlap; j.first = false;
                jobs[i].push_back(std::move(j));
            }
        }
        rounds = std::max(rounds, jobs[i].size());
    }
    for (size_t k = 0; k < rounds; ++k) {
        std::vector<Job *> round;
        std::vector<size_t> owner;
        for (size_t i = 0; i < active.size(); ++i)
            if (k < jobs[i].size() && rcs[i] == AMIRA_OK) { round.push_back(&jobs[i][k]); owner.push_back(i); }  // `?`: a failed session stops
        if (round.empty()) continue;
        run_round(g, round);
        for (size_t r = 0; r < round.size(); ++r) {
            Job &j = *round[r];
            Session &s = *j.s;
            if (j.rc) { rcs[owner[r]] = j.rc; continue; }
            if (j.first) { s.token_ids = j.tokens; s.transcript = decode_utf8(j.text.c_str()); }
            else accumulate(s, j);
        }
    }
    return AMIRA_OK;
}

// ---- incremental mode (amira_b200.h: amira_stream_group_set_incremental) ----
constexpr int64_t kHopS = 160, kHalfWin = 200, kCtxHops = 2;  // a frame's window covers +-200 samples around its centre 160 t

// the decoder states that live in the group's batch-layout buffers go back to their sessions
void flush_resident(amira_stream_group *g) {
    const size_t B = g->resident.size(), H = AMIRA_STATE_SIZE;
    if (B && g->resident_on_device) {
        amira_ctx_synchronize(g->p->ctx);
        const bool ok = cudaMemcpy(g->is1.data(), g->ds1.p, sizeof(float) * 2 * B * H, cudaMemcpyDeviceToHost) == cudaSuccess &&
                        cudaMemcpy(g->is2.data(), g->ds2.p, sizeof(float) * 2 * B * H, cudaMemcpyDeviceToHost) == cudaSuccess;
        g->resident_on_device = false;
        if (!ok) { cudaGetLastError(); g->resident.clear(); return; }  // a dead device: the sessions keep the states of their last flush
    }
    for (size_t i = 0; i < B; ++i)
        for (size_t l = 0; l < 2; ++l) {
            std::memcpy(g->resident[i]->s1.data() + l * H, g->is1.data() + (l * B + i) * H, sizeof(float) * H);
            std::memcpy(g->resident[i]->s2.data() + l * H, g->is2.data() + (l * B + i) * H, sizeof(float) * H);
        }
    g->resident.clear();
}

void refresh_transcript(amira_stream_group *g, Session &s) {
    if (!s.transcript_stale) return;
    s.transcript = decode_utf8(g->p->vocab.decode(s.token_ids.data(), (int32_t)s.token_ids.size()).c_str());
    s.transcript_stale = false;
}

// Frames [t_next, t_end) of every listed session, in one front-end launch, the injected encoder and one resumed decode launch.
// A session's segment starts at sample 160 A, A = max(0, t_next - 2): local frame t' = t - A.  For A > 0 the local frames 0 and 1
// see the segment's artificial left edge (reflect padding, missing pre-emphasis partner) and are dropped — t_next - A = 2 is the
// first local frame whose 400 window samples and their predecessors are all real.  For A = 0 the left edge is the true start of
// the stream.  The right edge is real audio (frames are only emitted once complete) or, on flush, the true end of the stream.
int32_t incremental_round(amira_stream_group *g, const std::vector<Session *> &ss, const std::vector<int64_t> &t_end, std::vector<int32_t> &rcs) {
    amira_pipeline *p = g->p;
    std::lock_guard<std::mutex> plock(p->mu);
    const int B = (int)ss.size();
    if (B == 0) return AMIRA_OK;
    static const bool sg_trace = getenv("AMIRA_SG_TRACE") != nullptr;  // debug: wall clock of the phases of a round on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_0 = now();
    double t_ph[6] = {0, 0, 0, 0, 0, 0};
    int ph = 0;
    auto lap = [&]() { const auto t1 = now(); t_ph[ph++] = std::chrono::duration<double, std::milli>(t1 - t_0).count(); t_0 = t1; };
    g->n_rounds++;
    g->n_pipeline_calls += B;
    std::vector<int64_t> A((size_t)B), nloc((size_t)B), lloc((size_t)B);
    g->woff.assign((size_t)B + 1, 0);
    g->foff.assign((size_t)B + 1, 0);
    for (int i = 0; i < B; ++i) {
        Session &s = *ss[(size_t)i];
        A[(size_t)i] = std::max<int64_t>(0, s.t_next - kCtxHops);
        nloc[(size_t)i] = s.n_total - kHopS * A[(size_t)i];
        lloc[(size_t)i] = nloc[(size_t)i] / kHopS + 1;
        g->woff[(size_t)i + 1] = g->woff[(size_t)i] + nloc[(size_t)i];
        g->foff[(size_t)i + 1] = g->foff[(size_t)i] + (int64_t)AMIRA_N_MELS * lloc[(size_t)i];
    }
    g->pcm.resize((size_t)std::max<int64_t>(g->woff[(size_t)B], 1));
    parallel_for(B, 128, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            const Session &s = *ss[(size_t)i];
            const int64_t a = kHopS * A[(size_t)i] - s.hist_start;  // >= 0: the history keeps two hops of left context
            std::memcpy(g->pcm.data() + g->woff[(size_t)i], s.hist.data() + a, sizeof(int16_t) * (size_t)nloc[(size_t)i]);
        }
    });
    g->ifeat.resize((size_t)std::max<int64_t>(g->foff[(size_t)B], 1));
    g->flens.assign((size_t)B, 0);
    lap();  // 0: segments gathered
    int32_t rc = amira_logmel_pcm16_packed(p->ctx, g->pcm.data(), g->woff.data(), B, g->ifeat.data(), g->foff.data(), g->flens.data());
    if (rc) { g->err = amira_last_error(p->ctx); for (auto &r : rcs) r = rc; return rc; }
    lap();  // 1: front end
    // running normalisation of the new frames, sessions in parallel: Chan merge of the new frames into the running (count, mean,
    // M2) of each feature, then (x - mean) / (std + 1e-5) with the merged statistics (sample variance, like the utterance-level
    // rule of the preprocessor)
    g->coff.assign((size_t)B + 1, 0);
    for (int i = 0; i < B; ++i) g->coff[(size_t)i + 1] = g->coff[(size_t)i] + (int64_t)AMIRA_N_MELS * (t_end[(size_t)i] - ss[(size_t)i]->t_next);
    g->ichunks.resize((size_t)std::max<int64_t>(g->coff[(size_t)B], 1));
    parallel_for(B, 32, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            Session &s = *ss[(size_t)i];
            const int64_t n_new = t_end[(size_t)i] - s.t_next, f_first = s.t_next - A[(size_t)i], L = lloc[(size_t)i];
            const float *src = g->ifeat.data() + g->foff[(size_t)i];  // [128][L]
            float *chunk = g->ichunks.data() + g->coff[(size_t)i];   // [128][n_new]
            for (int m = 0; m < AMIRA_N_MELS; ++m) {
                const float *row = src + (size_t)m * L + f_first;
                double bm = 0.0, b2 = 0.0;
                for (int64_t f = 0; f < n_new; ++f) bm += (double)row[f];
                bm /= (double)n_new;
                for (int64_t f = 0; f < n_new; ++f) { const double d = (double)row[f] - bm; b2 += d * d; }
                const double na = (double)s.st_count, nb = (double)n_new, nt = na + nb, d = bm - s.st_mean[(size_t)m];
                const double mean = s.st_mean[(size_t)m] + d * (nb / nt), m2 = s.st_m2[(size_t)m] + b2 + d * d * (na * nb / nt);
                s.st_mean[(size_t)m] = mean;
                s.st_m2[(size_t)m] = m2;
                const double sd = nt > 1.0 ? std::sqrt(m2 / (nt - 1.0)) : 0.0;
                const float mu = (float)mean, inv = (float)(1.0 / (sd + 1e-5));
                for (int64_t f = 0; f < n_new; ++f) chunk[(size_t)m * n_new + f] = (row[f] - mu) * inv;
            }
            s.st_count += n_new;
            s.t_next = t_end[(size_t)i];
        }
    });
    // the injected encoder on the new frames, one stream at a time (its output buffer is only valid until its next call)
    g->eoff.assign((size_t)B + 1, 0);
    g->elens.assign((size_t)B, 0);
    for (int i = 0; i < B; ++i) {
        const int64_t n_new = (g->coff[(size_t)i + 1] - g->coff[(size_t)i]) / AMIRA_N_MELS;
        const float *e = nullptr;
        int64_t el = 0;
        if (!p->encoder || p->encoder(p->encoder_user, g->ichunks.data() + g->coff[(size_t)i], n_new, &e, &el) != 0 || el < 0 || (el > 0 && !e)) {
            rcs[(size_t)i] = p->encoder ? AMIRA_ERR_UNKNOWN : AMIRA_ERR_NOT_READY;
            el = 0;
        }
        g->eoff[(size_t)i + 1] = g->eoff[(size_t)i] + (int64_t)AMIRA_ENC_DIM * el;
        if (el > 0) {
            g->ienc.resize((size_t)g->eoff[(size_t)i + 1]);
            std::memcpy(g->ienc.data() + g->eoff[(size_t)i], e, sizeof(float) * (size_t)AMIRA_ENC_DIM * (size_t)el);
        }
        g->elens[(size_t)i] = el;
    }
    lap();  // 2: running normalisation + encoder
    // resumed greedy loop: state and last token in and out.  The states of the round's streams stay in the group's batch-layout
    // buffers from round to round while the list of streams does not change.
    int32_t cap = AMIRA_MAX_TOTAL_TOKENS;
    amira_ctx_max_total_tokens(p->ctx, &cap);
    const size_t H = AMIRA_STATE_SIZE;
    if (g->resident != ss) {
        flush_resident(g);
        g->is1.resize(2 * (size_t)B * H);
        g->is2.resize(2 * (size_t)B * H);
        parallel_for(B, 128, [&](int lo, int hi) {
            for (int i = lo; i < hi; ++i)
                for (int l = 0; l < 2; ++l) {
                    std::memcpy(g->is1.data() + ((size_t)l * B + i) * H, ss[(size_t)i]->s1.data() + (size_t)l * H, sizeof(float) * H);
                    std::memcpy(g->is2.data() + ((size_t)l * B + i) * H, ss[(size_t)i]->s2.data() + (size_t)l * H, sizeof(float) * H);
                }
        });
        g->resident = ss;
        g->resident_on_device = g->ds1.reserve(2 * (size_t)B * H) && g->ds2.reserve(2 * (size_t)B * H) &&
                                cudaMemcpy(g->ds1.p, g->is1.data(), sizeof(float) * 2 * (size_t)B * H, cudaMemcpyHostToDevice) == cudaSuccess &&
                                cudaMemcpy(g->ds2.p, g->is2.data(), sizeof(float) * 2 * (size_t)B * H, cudaMemcpyHostToDevice) == cudaSuccess;
        if (!g->resident_on_device) cudaGetLastError();
    }
    g->last.resize((size_t)B);
    for (int i = 0; i < B; ++i) g->last.data()[i] = ss[(size_t)i]->last_token;
    g->itokens.resize((size_t)B * (size_t)cap);
    g->intok.resize((size_t)B);
    std::memset(g->intok.data(), 0, sizeof(int32_t) * (size_t)B);
    lap();  // 3: states gathered
    if (g->eoff[(size_t)B] > 0) {
        rc = amira_greedy_decode_resume(p->ctx, g->ienc.data(), g->eoff.data(), B, g->elens.data(), g->resident_on_device ? g->ds1.p : g->is1.data(),
                                        g->resident_on_device ? g->ds2.p : g->is2.data(), g->last.data(), g->itokens.data(), g->intok.data(), nullptr);
        if (rc && rc != AMIRA_ERR_DECODE_STEP) {
            g->resident.clear();  // the buffers may hold a partial result: the sessions keep the states of their last flush
            g->resident_on_device = false;
            g->err = amira_last_error(p->ctx);
            for (auto &r : rcs) r = rc;
            return rc;
        }
    }
    lap();  // 4: decode
    for (int i = 0; i < B; ++i) {
        Session &s = *ss[(size_t)i];
        if (rcs[(size_t)i]) continue;
        const int32_t nt = g->intok.data()[i];
        if (nt < 0) { rcs[(size_t)i] = AMIRA_ERR_DECODE_STEP; continue; }
        if (g->elens[(size_t)i] > 0) {
            s.last_token = g->last.data()[i];
            s.enc_frames += g->elens[(size_t)i];
        }
        if (nt > 0) {
            const int32_t *tk = g->itokens.data() + (size_t)i * (size_t)cap;
            s.token_ids.insert(s.token_ids.end(), tk, tk + nt);
            s.transcript_stale = true;  // rebuilt on demand (amira_stream_group_transcript): decoding every stream's whole history
                                        // on every tick cost more than the GPU work of the tick
        }
    }
    lap();  // 5: results scattered
    if (sg_trace)
        fprintf(stderr, "[stream group] B=%d gather %.2f | front end %.2f | normalise+encoder %.2f | states %.2f | decode %.2f | scatter %.2f ms"
                " (pcm %.1f MB %s, features %.1f MB %s, encoder outputs %.1f MB %s)\n", B,
                t_ph[0], t_ph[1], t_ph[2], t_ph[3], t_ph[4], t_ph[5], g->pcm.size() * 2e-6, g->pcm.pinned ? "pinned" : "pageable",
                g->ifeat.size() * 4e-6, g->ifeat.pinned ? "pinned" : "pageable", g->ienc.size() * 4e-6, g->ienc.pinned ? "pinned" : "pageable");
    return AMIRA_OK;
}

// push the new audio of the listed streams, then emit whatever has become complete (or, on flush, everything that is left)
int32_t incremental_step(amira_stream_group *g, int32_t n, const int32_t *streams, const uint8_t *const *audio_bytes, const size_t *n_bytes,
                         bool flush, int32_t *status) {
    std::vector<Session *> ss;
    std::vector<int64_t> t_end;
    std::vector<int32_t> idx;
    for (int32_t i = 0; i < n; ++i) {
        Session &s = *g->sessions[(size_t)streams[i]];
        if (status) status[i] = AMIRA_OK;
        if (s.flushed) { if (status) status[i] = AMIRA_ERR_INVALID_VALUE; continue; }  // a flushed stream takes no more audio (clear it first)
        if (!flush && n_bytes[i]) {
            const size_t ns = n_bytes[i] / 2;  // an odd trailing byte is dropped (audio.rs:18-26)
            const size_t o = s.hist.size();
            s.hist.resize(o + ns);
            std::memcpy(s.hist.data() + o, audio_bytes[i], 2 * ns);  // little-endian i16, as the wire carries it
            s.n_total += (int64_t)ns;
        }
        int64_t te;
        if (flush) {
            s.flushed = true;
            te = s.n_total > 0 ? s.n_total / kHopS + 1 : 0;  // features_lens of the whole stream
        } else {
            te = s.n_total >= kHalfWin ? (s.n_total - kHalfWin) / kHopS + 1 : 0;  // frames whose window lies inside the audio so far
        }
        if (te > s.t_next) { ss.push_back(&s); t_end.push_back(te); idx.push_back(i); }
    }
    std::vector<int32_t> rcs(ss.size(), AMIRA_OK);
    incremental_round(g, ss, t_end, rcs);
    int32_t worst = AMIRA_OK;
    for (size_t a = 0; a < ss.size(); ++a) {
        Session &s = *ss[a];
        // keep two hops of left context before the next frame (and one more sample: the pre-emphasis partner lies inside them)
        const int64_t keep_from = std::max<int64_t>(0, kHopS * (s.t_next - kCtxHops));
        if (keep_from > s.hist_start) {
            s.hist.erase(s.hist.begin(), s.hist.begin() + (keep_from - s.hist_start));
            s.hist_start = keep_from;
        }
        if (status) status[idx[a]] = rcs[a];
        if (rcs[a] && !worst) worst = rcs[a];
    }
    if (status)
        for (int32_t i = 0; i < n; ++i)
            if (status[i] && !worst) worst = status[i];
    return worst;
}

void copy_text(const std::string &s, char *text, size_t cap, int32_t *len) {
    if (len) *len = (int32_t)s.size();
    if (text && cap) {
        const size_t m = std::min(s.size(), cap - 1);
        std::memcpy(text, s.data(), m);
        text[m] = '\0';
    }
}

}  // namespace

// no exception crosses the C boundary (std::bad_alloc / std::length_error from the string and vector work below)
#define HS_TRY try {
#define HS_CATCH(g_)                                                                                            \
    } catch (const std::bad_alloc &) { return gfail((g_), AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed"); } \
    catch (const std::exception &ex) { return gfail((g_), AMIRA_ERR_UNKNOWN, ex.what()); }                     \
    catch (...) { return gfail((g_), AMIRA_ERR_UNKNOWN, "unexpected exception"); }

extern "C" {

int32_t amira_weave_transcript_segs(const char *first_seg, const char *second_seg, float percent_time_overlap, float min_alignment_score,
                                    char *out, size_t out_cap, int32_t *out_len) {
    if (!first_seg || !second_seg) return AMIRA_ERR_INVALID_VALUE;
    HS_TRY
    copy_text(encode_utf8(weave_transcript_segs(decode_utf8(first_seg), decode_utf8(second_seg), percent_time_overlap, min_alignment_score)),
              out, out_cap, out_len);
    return AMIRA_OK;
    HS_CATCH(nullptr)
}

int32_t amira_best_alignment(const char *first, const char *second, float percent_time_overlap, int32_t *overlap, float *score) {
    if (!first || !second || !overlap || !score) return AMIRA_ERR_INVALID_VALUE;
    HS_TRY
    const u32s a = decode_utf8(first), b = decode_utf8(second);
    size_t o;
    best_alignment({a.data(), a.size()}, {b.data(), b.size()}, percent_time_overlap, &o, score);
    *overlap = (int32_t)o;
    return AMIRA_OK;
    HS_CATCH(nullptr)
}

int32_t amira_is_overlap_silence(const float *overlap_audio, size_t n, float mean_amplitude, int32_t *silent) {
    if (!silent || (n && !overlap_audio)) return AMIRA_ERR_INVALID_VALUE;
    *silent = is_overlap_silence(overlap_audio, n, mean_amplitude) ? 1 : 0;
    return AMIRA_OK;
}

int32_t amira_mean_amplitude(const float *samples, size_t n, float *mean) {
    if (!mean || (n && !samples)) return AMIRA_ERR_INVALID_VALUE;
    *mean = mean_amplitude_of(samples, n);
    return AMIRA_OK;
}

int32_t amira_window_sequence(int64_t total_len, int64_t window_size, int64_t leading_context, int64_t trailing_context, int64_t *slices,
                              float *overlap_ratio, int32_t cap, int32_t *n_windows) {
    if (total_len < 0 || window_size <= 0 || leading_context < 0 || trailing_context < 0 || !n_windows ||
        window_size <= leading_context + trailing_context)
        return AMIRA_ERR_INVALID_VALUE;
    HS_TRY
    const std::vector<Window> w = window_sequence((size_t)total_len, (size_t)window_size, (size_t)leading_context, (size_t)trailing_context);
    *n_windows = (int32_t)w.size();
    for (int32_t i = 0; i < *n_windows && i < cap; ++i) {
        if (slices) {
            slices[4 * i + 0] = (int64_t)w[(size_t)i].src_start; slices[4 * i + 1] = (int64_t)w[(size_t)i].src_end;
            slices[4 * i + 2] = (int64_t)w[(size_t)i].tgt_start; slices[4 * i + 3] = (int64_t)w[(size_t)i].tgt_end;
        }
        if (overlap_ratio) overlap_ratio[i] = w[(size_t)i].overlap;
    }
    return AMIRA_OK;
    HS_CATCH(nullptr)
}

int32_t amira_stream_group_create(amira_pipeline *p, int32_t n_streams, float chunk_size, float leading_context, float trailing_context,
                                  float buffer_capacity, amira_stream_group **out) {
    if (!p || !out || n_streams <= 0 || !(chunk_size > 0.f) || leading_context < 0.f || trailing_context < 0.f || !(buffer_capacity > 0.f) ||
        f32_as_usize(chunk_size * 16000.0f) == 0)  // a zero-sample chunk would never advance window_sequence
        return AMIRA_ERR_INVALID_VALUE;
    auto *g = new (std::nothrow) amira_stream_group();
    if (!g) return AMIRA_ERR_OUT_OF_MEMORY;
    g->p = p;
    try {
        for (int32_t i = 0; i < n_streams; ++i) {
            std::unique_ptr<Session> s(new Session());
            s->audio.init(f32_as_usize(buffer_capacity * 16000.0f), chunk_size, leading_context, trailing_context);  // incremental.rs:80-90
            s->chunk_size = chunk_size;
            s->reset_state();
            g->sessions.push_back(std::move(s));
        }
    } catch (...) {
        delete g;
        return AMIRA_ERR_OUT_OF_MEMORY;
    }
    *out = g;
    return AMIRA_OK;
}

int32_t amira_stream_group_destroy(amira_stream_group *g) {
    delete g;
    return AMIRA_OK;
}

const char *amira_stream_group_last_error(amira_stream_group *g) { return g ? g->err.c_str() : "null stream group"; }

int32_t amira_stream_group_clear(amira_stream_group *g, int32_t stream) {
    if (!g || stream < 0 || (size_t)stream >= g->sessions.size()) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    flush_resident(g);  // the other streams' decoder states go back to their sessions before this one is reset
    g->sessions[(size_t)stream]->clear();
    return AMIRA_OK;
}

// IncrementalAsr::process_chunk (incremental.rs:111-129) for n distinct streams at once
int32_t amira_stream_group_process_chunks(amira_stream_group *g, int32_t n, const int32_t *streams, const uint8_t *const *audio_bytes,
                                          const size_t *n_bytes, int32_t *status) {
    if (!g) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    if (n < 0 || (n > 0 && (!streams || !audio_bytes || !n_bytes))) return gfail(g, AMIRA_ERR_INVALID_VALUE, "process_chunks: bad arguments");
    if (!g->p->ctx) return gfail(g, AMIRA_ERR_NO_DEVICE, "pipeline was created without a GPU context");
    HS_TRY
        std::vector<char> seen(g->sessions.size(), 0);
        for (int32_t i = 0; i < n; ++i) {
            if (streams[i] < 0 || (size_t)streams[i] >= g->sessions.size() || seen[(size_t)streams[i]])
                return gfail(g, AMIRA_ERR_INVALID_VALUE, "process_chunks: stream id out of range or repeated");
            seen[(size_t)streams[i]] = 1;
            if (n_bytes[i] && !audio_bytes[i]) return gfail(g, AMIRA_ERR_INVALID_VALUE, "process_chunks: null audio");
        }
        if (g->incremental) return incremental_step(g, n, streams, audio_bytes, n_bytes, false, status);
        std::vector<Session *> active;
        std::vector<int32_t> idx;
        std::vector<float> samples;
        for (int32_t i = 0; i < n; ++i) {
            Session &s = *g->sessions[(size_t)streams[i]];
            const size_t ns = n_bytes[i] / 2;  // bytes_to_f32_samples (audio.rs:18-26): an odd trailing byte is dropped
            samples.resize(ns);
            for (size_t k = 0; k < ns; ++k) {
                const int16_t v = (int16_t)((uint16_t)audio_bytes[i][2 * k] | ((uint16_t)audio_bytes[i][2 * k + 1] << 8));
                samples[k] = (float)v / 32768.0f;
            }
            s.audio.add_samples(samples.data(), ns);
            s.acc_mean_amplitude = s.audio.mean_amplitude;
            if (status) status[i] = AMIRA_OK;
            if (s.audio.length != 0) { active.push_back(&s); idx.push_back(i); }
        }
        std::vector<int32_t> rcs(active.size(), AMIRA_OK);
        process_buffered(g, active, rcs);
        int32_t worst = AMIRA_OK;
        for (size_t a = 0; a < active.size(); ++a) {
            if (status) status[idx[a]] = rcs[a];
            if (rcs[a] && !worst) worst = rcs[a];
        }
        return worst;
    HS_CATCH(g)
}

int32_t amira_stream_group_transcript(amira_stream_group *g, int32_t stream, char *text, size_t text_cap, int32_t *text_len) {
    if (!g || stream < 0 || (size_t)stream >= g->sessions.size()) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    HS_TRY
    refresh_transcript(g, *g->sessions[(size_t)stream]);
    copy_text(encode_utf8(g->sessions[(size_t)stream]->transcript), text, text_cap, text_len);
    return AMIRA_OK;
    HS_CATCH(g)
}

int32_t amira_stream_group_tokens(amira_stream_group *g, int32_t stream, int32_t *tokens, int32_t tokens_cap, int32_t *n_tokens) {
    if (!g || stream < 0 || (size_t)stream >= g->sessions.size() || !n_tokens) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    const std::vector<int32_t> &t = g->sessions[(size_t)stream]->token_ids;
    *n_tokens = (int32_t)t.size();
    if (tokens && tokens_cap > 0) std::memcpy(tokens, t.data(), sizeof(int32_t) * std::min<size_t>(t.size(), (size_t)tokens_cap));
    return AMIRA_OK;
}

int32_t amira_stream_group_audio_length(amira_stream_group *g, int32_t stream, float *seconds) {  // incremental.rs:295-297
    if (!g || stream < 0 || (size_t)stream >= g->sessions.size() || !seconds) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    *seconds = (float)g->sessions[(size_t)stream]->audio.length / 16000.0f;
    return AMIRA_OK;
}

// IncrementalAsr::process_batch (incremental.rs:267-292)
int32_t amira_stream_group_process_batch(amira_stream_group *g, int32_t stream, const uint8_t *audio_bytes, size_t n_bytes,
                                         amira_transcription *out, int32_t *tokens, int32_t tokens_cap, char *text, size_t text_cap) {
    if (!g || stream < 0 || (size_t)stream >= g->sessions.size() || !out || (n_bytes && !audio_bytes)) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    flush_resident(g);
    Session &s = *g->sessions[(size_t)stream];
    s.clear();
    const size_t ns = n_bytes / 2;
    if ((float)ns / 16000.0f <= s.chunk_size) {
        const int32_t rc = amira_pipeline_process_batch(g->p, audio_bytes, n_bytes, out, tokens, tokens_cap, text, text_cap);
        if (rc) g->err = amira_pipeline_last_error(g->p);
        return rc;
    }
    HS_TRY
        std::vector<float> samples(ns);
        for (size_t k = 0; k < ns; ++k) {
            const int16_t v = (int16_t)((uint16_t)audio_bytes[2 * k] | ((uint16_t)audio_bytes[2 * k + 1] << 8));
            samples[k] = (float)v / 32768.0f;
        }
        s.audio.add_samples(samples.data(), ns);
        std::vector<Session *> active{&s};
        std::vector<int32_t> rcs(1, AMIRA_OK);
        process_buffered(g, active, rcs);
        if (rcs[0]) return rcs[0];
        std::memset(out, 0, sizeof(*out));
        out->audio_length_samples = (int64_t)ns;  // features_length / encoded_length stay 0 (incremental.rs:288-289)
        out->n_tokens = (int32_t)s.token_ids.size();
        if (tokens && tokens_cap > 0) std::memcpy(tokens, s.token_ids.data(), sizeof(int32_t) * std::min<size_t>(s.token_ids.size(), (size_t)tokens_cap));
        copy_text(encode_utf8(s.transcript), text, text_cap, &out->text_len);
        return AMIRA_OK;
    HS_CATCH(g)
}

int32_t amira_stream_group_set_incremental(amira_stream_group *g, int32_t enable) {
    if (!g) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    for (auto &s : g->sessions)
        if (s->audio.length != 0 || s->n_total != 0) return gfail(g, AMIRA_ERR_INVALID_VALUE, "set_incremental: a stream already holds audio (clear it first)");
    flush_resident(g);
    g->incremental = enable != 0;
    return AMIRA_OK;
}

int32_t amira_stream_group_flush(amira_stream_group *g, int32_t n, const int32_t *streams, int32_t *status) {
    if (!g) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    if (!g->incremental) return gfail(g, AMIRA_ERR_INVALID_VALUE, "flush: the group is not in incremental mode");
    if (n < 0 || (n > 0 && !streams)) return gfail(g, AMIRA_ERR_INVALID_VALUE, "flush: bad arguments");
    if (!g->p->ctx) return gfail(g, AMIRA_ERR_NO_DEVICE, "pipeline was created without a GPU context");
    HS_TRY
        std::vector<char> seen(g->sessions.size(), 0);
        for (int32_t i = 0; i < n; ++i) {
            if (streams[i] < 0 || (size_t)streams[i] >= g->sessions.size() || seen[(size_t)streams[i]])
                return gfail(g, AMIRA_ERR_INVALID_VALUE, "flush: stream id out of range or repeated");
            seen[(size_t)streams[i]] = 1;
        }
        return incremental_step(g, n, streams, nullptr, nullptr, true, status);
    HS_CATCH(g)
}

int32_t amira_stream_group_progress(amira_stream_group *g, int32_t stream, int64_t *samples, int64_t *frames, int64_t *encoder_frames) {
    if (!g || stream < 0 || (size_t)stream >= g->sessions.size()) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    const Session &s = *g->sessions[(size_t)stream];
    if (samples) *samples = s.n_total;
    if (frames) *frames = s.t_next;
    if (encoder_frames) *encoder_frames = s.enc_frames;
    return AMIRA_OK;
}

int32_t amira_stream_group_stats(amira_stream_group *g, int64_t *n_pipeline_calls, int64_t *n_rounds) {
    if (!g) return AMIRA_ERR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g->mu);
    if (n_pipeline_calls) *n_pipeline_calls = g->n_pipeline_calls;
    if (n_rounds) *n_rounds = g->n_rounds;
    return AMIRA_OK;
}

}  // extern "C"
