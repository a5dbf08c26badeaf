// host_pipeline.cpp — C++ host side above the C ABI: the mirror of the reference's Rust `AsrPipeline` trait
// (src/asr/pipeline.rs:20-67) and of TritonAsrPipeline::process_audio_zero_copy (src/asr/pipeline.rs:269-380),
// with `preprocessor` and `decoder_joint` served by the GPU library instead of Triton.  Rust is not available in
// this build environment; INTEGRATION.md carries the equivalent Rust binding as source.
// Also: Vocabulary (src/asr/types.rs:77-155) and the host-side utterance sharder for multi-GPU.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <thread>
#include <cstring>
#include <fstream>
#include <mutex>
#include <numeric>
#include <queue>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "amira_b200.h"

#include "host_common.h"

using amira_host::Vocabulary;

namespace {

thread_local std::string g_err;

int32_t pfail(amira_pipeline *p, int32_t code, const std::string &m) {
    if (p) p->err = m; else g_err = m;
    return code;
}

// process_audio_zero_copy (src/asr/pipeline.rs:269-380) for one request
int32_t run_request(amira_pipeline *p, const uint8_t *bytes, size_t n_bytes, const float *samples, size_t n_samples,
                    float *states_1, float *states_2, amira_transcription *out, int32_t *tokens, int32_t tokens_cap,
                    char *text, size_t text_cap) {
    if (!p) return pfail(nullptr, AMIRA_ERR_INVALID_VALUE, "null pipeline");
    std::lock_guard<std::mutex> lock(p->mu);
    if (!out) return pfail(p, AMIRA_ERR_INVALID_VALUE, "null transcription");
    if (tokens_cap < 0) return pfail(p, AMIRA_ERR_INVALID_VALUE, "negative tokens_cap");
    std::memset(out, 0, sizeof(*out));
    if (!p->ctx) return pfail(p, AMIRA_ERR_NO_DEVICE, "pipeline was created without a GPU context (vocabulary only)");
    if (text && text_cap) text[0] = '\0';
    int64_t n = 0, flen = 0;
    int32_t rc;
    // step 0/1: convert_audio (:127-139) + preprocessor (:283-291), fused on the GPU
    if (bytes && (n_bytes % 2 == 0)) {
        n = (int64_t)(n_bytes / 2);
        const int64_t offs[2] = {0, n};
        amira_features_len(n, &flen);
        p->features.resize((size_t)AMIRA_N_MELS * (size_t)std::max<int64_t>(flen, 1));
        rc = amira_preprocess_pcm16(p->ctx, reinterpret_cast<const int16_t *>(bytes), offs, 1, p->features.data(),
                                    std::max<int64_t>(flen, 1), &flen);
    } else {
        if (bytes) {  // odd length: bytes_to_f32_optimized's trailing-byte rule (src/performance_opts.rs:26-30)
            size_t got = 0;
            p->wave.resize(n_bytes / 2 + 1);
            rc = amira_bytes_to_f32(p->ctx, bytes, n_bytes, 0, p->wave.data(), &got);
            if (rc) return pfail(p, rc, amira_last_error(p->ctx));
            samples = p->wave.data();
            n_samples = got;
        }
        n = (int64_t)n_samples;
        amira_features_len(n, &flen);
        p->features.resize((size_t)AMIRA_N_MELS * (size_t)std::max<int64_t>(flen, 1));
        rc = amira_preprocess_f32(p->ctx, samples, n, &n, 1, p->features.data(), std::max<int64_t>(flen, 1), &flen);
    }
    if (rc) return pfail(p, rc, amira_last_error(p->ctx));
    out->audio_length_samples = n;
    out->features_length = flen;
    // step 2: encoder (out of scope; injected)
    if (!p->encoder) return pfail(p, AMIRA_ERR_NOT_READY, "no encoder callback installed");
    const float *enc = nullptr;
    int64_t enc_len = 0;
    if (p->encoder(p->encoder_user, p->features.data(), flen, &enc, &enc_len) != 0 || enc_len < 0 || (enc_len > 0 && !enc))
        return pfail(p, AMIRA_ERR_UNKNOWN, "encoder callback failed");
    out->encoded_length = enc_len;
    // step 3: greedy decode (:313-356) — one persistent kernel instead of one RPC per step
    int32_t row_cap = AMIRA_MAX_TOTAL_TOKENS;  // the decode entry writes max_total_tokens ids per stream: size the row from the context
    if ((rc = amira_ctx_max_total_tokens(p->ctx, &row_cap)) != 0 || row_cap <= 0) return pfail(p, rc ? rc : AMIRA_ERR_UNKNOWN, "cannot query max_total_tokens");
    p->tokens.assign((size_t)row_cap, 0);
    int32_t ntok = 0;
    if (enc_len > 0) {
        rc = amira_greedy_decode(p->ctx, enc, 1, (int32_t)enc_len, &enc_len, states_1, states_2, p->tokens.data(), &ntok, nullptr);
        if (rc) return pfail(p, rc, amira_last_error(p->ctx));
    }
    ntok = std::max(0, std::min(ntok, row_cap));
    out->n_tokens = ntok;
    if (tokens) std::memcpy(tokens, p->tokens.data(), sizeof(int32_t) * (size_t)std::min(ntok, tokens_cap));
    // step 4: tokens -> text (:361-363)
    const std::string s = p->vocab.decode(p->tokens.data(), ntok);
    out->text_len = (int32_t)s.size();
    if (text && text_cap) {
        const size_t m = std::min(s.size(), text_cap - 1);
        std::memcpy(text, s.data(), m);
        text[m] = '\0';
    }
    return AMIRA_OK;
}

}  // namespace

extern "C" {

int32_t amira_pipeline_create(amira_ctx *ctx, const char *vocab_path, amira_encoder_fn encoder, void *encoder_user,
                              amira_pipeline **out) {
    if (!out) return pfail(nullptr, AMIRA_ERR_INVALID_VALUE, "null argument");  // ctx may be NULL: vocabulary-only use
    *out = nullptr;
    amira_pipeline *p = nullptr;
    try {
        p = new amira_pipeline();
        p->ctx = ctx;
        p->encoder = encoder;
        p->encoder_user = encoder_user;
        if (vocab_path && !p->vocab.load(vocab_path)) {
            delete p;
            return pfail(nullptr, AMIRA_ERR_IO, std::string("cannot read vocabulary ") + vocab_path);
        }
    } catch (...) {
        delete p;
        return pfail(nullptr, AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    *out = p;
    return AMIRA_OK;
}

int32_t amira_pipeline_destroy(amira_pipeline *p) {
    delete p;
    return AMIRA_OK;
}

const char *amira_pipeline_last_error(amira_pipeline *p) { return p ? p->err.c_str() : g_err.c_str(); }

#define GUARD(expr)                                                                       \
    try {                                                                                 \
        return (expr);                                                                    \
    } catch (const std::bad_alloc &) {                                                    \
        return pfail(p, AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");               \
    } catch (...) {                                                                       \
        return pfail(p, AMIRA_ERR_UNKNOWN, "unexpected exception");                       \
    }

int32_t amira_pipeline_process_batch(amira_pipeline *p, const uint8_t *audio_bytes, size_t n_bytes, amira_transcription *out,
                                     int32_t *tokens, int32_t tokens_cap, char *text, size_t text_cap) {
    static const uint8_t kEmpty = 0;
    GUARD(run_request(p, audio_bytes ? audio_bytes : &kEmpty, audio_bytes ? n_bytes : 0, nullptr, 0, nullptr, nullptr, out,
                      tokens, tokens_cap, text, text_cap))
}

int32_t amira_pipeline_process_stream_chunk(amira_pipeline *p, const uint8_t *audio_bytes, size_t n_bytes, float *states_1,
                                            float *states_2, amira_transcription *out, int32_t *tokens,
                                            int32_t tokens_cap, char *text, size_t text_cap) {
    static const uint8_t kEmpty = 0;
    if (p && (!states_1 || !states_2)) return pfail(p, AMIRA_ERR_INVALID_VALUE, "stream calls need both state buffers");
    GUARD(run_request(p, audio_bytes ? audio_bytes : &kEmpty, audio_bytes ? n_bytes : 0, nullptr, 0, states_1, states_2, out,
                      tokens, tokens_cap, text, text_cap))
}

int32_t amira_pipeline_process_batch_samples(amira_pipeline *p, const float *samples, size_t n_samples,
                                             amira_transcription *out, int32_t *tokens, int32_t tokens_cap, char *text,
                                             size_t text_cap) {
    static const float kEmpty = 0.f;
    GUARD(run_request(p, nullptr, 0, samples ? samples : &kEmpty, samples ? n_samples : 0, nullptr, nullptr, out, tokens,
                      tokens_cap, text, text_cap))
}

int32_t amira_pipeline_process_stream_samples(amira_pipeline *p, const float *samples, size_t n_samples, float *states_1,
                                              float *states_2, amira_transcription *out, int32_t *tokens,
                                              int32_t tokens_cap, char *text, size_t text_cap) {
    static const float kEmpty = 0.f;
    if (p && (!states_1 || !states_2)) return pfail(p, AMIRA_ERR_INVALID_VALUE, "stream calls need both state buffers");
    GUARD(run_request(p, nullptr, 0, samples ? samples : &kEmpty, samples ? n_samples : 0, states_1, states_2, out, tokens,
                      tokens_cap, text, text_cap))
}

int32_t amira_vocab_decode(amira_pipeline *p, const int32_t *tokens, int32_t n_tokens, char *text, size_t text_cap,
                           int32_t *text_len) {
    if (!p || (n_tokens > 0 && !tokens) || n_tokens < 0) return pfail(p, AMIRA_ERR_INVALID_VALUE, "bad arguments");
    try {
        const std::string s = p->vocab.decode(tokens, n_tokens);
        if (text_len) *text_len = (int32_t)s.size();
        if (text && text_cap) {
            const size_t m = std::min(s.size(), text_cap - 1);
            std::memcpy(text, s.data(), m);
            text[m] = '\0';
        }
    } catch (...) {
        return pfail(p, AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    return AMIRA_OK;
}

// Longest-processing-time-first bin assignment; ties broken by index so every rank computes the same map.
int32_t amira_shard_utterances(const int64_t *costs, int32_t n, int32_t n_shards, int32_t *shard_of) {
    if (n < 0 || n_shards <= 0 || (n > 0 && (!costs || !shard_of))) return AMIRA_ERR_INVALID_VALUE;
    try {
        std::vector<int32_t> order((size_t)n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return costs[a] > costs[b]; });
        using Load = std::pair<int64_t, int32_t>;  // (load, shard)
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
        for (int32_t s = 0; s < n_shards; ++s) heap.push({0, s});
        for (int32_t i : order) {
            Load l = heap.top();
            heap.pop();
            shard_of[i] = l.second;
            l.first += std::max<int64_t>(costs[i], 0);
            heap.push(l);
        }
    } catch (...) {
        return AMIRA_ERR_OUT_OF_MEMORY;
    }
    return AMIRA_OK;
}

}  // extern "C"


// =================================================================================================================
// Request micro-batcher.  The reference is strictly B = 1 per request (src/triton/model.rs:84,298,590)
// and bounds concurrency with semaphores (src/server/state.rs:47-61, 10 streams / 50 batches); the GPU path wants the
// opposite: concurrent `process_batch` calls coalesced into ONE front-end launch and ONE persistent decode launch.
// amira_batcher_process_batch is a blocking, thread-safe drop-in for AsrPipeline::process_batch: callers park on a
// condition variable while a worker thread drains the queue every `max_wait_us` (or as soon as `max_batch` requests
// wait), runs the batch and hands each caller its own Transcription.  Results are identical to the one-by-one calls:
// utterances are independent in every kernel.
namespace {

struct BatchReq {
    const uint8_t *bytes = nullptr;
    size_t n_bytes = 0;
    amira_transcription *out = nullptr;
    int32_t *tokens = nullptr;
    int32_t tokens_cap = 0;
    char *text = nullptr;
    size_t text_cap = 0;
    int32_t rc = 0;
    bool done = false;
    std::string err;
};

}  // namespace

struct amira_batcher {
    amira_pipeline *p = nullptr;
    int max_batch = 64;
    int max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<BatchReq *> queue;
    bool stop = false;
    std::thread worker;
    std::atomic<int64_t> n_requests{0}, n_batches{0};
    // worker-owned scratch
    std::vector<int16_t> pcm;
    std::vector<int64_t> offsets, foff, eoff, flens, elens;
    std::vector<float> features, enc;
    std::vector<int32_t> tokens, ntok;
};

namespace {

void finish_req(amira_batcher *b, BatchReq *r, int32_t rc, const std::string &err) {
    r->rc = rc;
    r->err = err;
    std::lock_guard<std::mutex> lock(b->mu);
    r->done = true;
    b->cv_done.notify_all();
}

// one coalesced batch: fused front end over all requests, the injected encoder per utterance, one greedy decode
void run_batch(amira_batcher *b, std::vector<BatchReq *> &reqs) {
    amira_pipeline *p = b->p;
    std::lock_guard<std::mutex> plock(p->mu);  // the pipeline's context calls are serialised like single requests
    const int B = (int)reqs.size();
    auto fail_all = [&](int32_t rc, const std::string &m) {
        // a finished request is nulled in `reqs` (its owner may destroy it the moment it is marked done): skip those
        for (BatchReq *&r : reqs)
            if (r) { finish_req(b, r, rc, m); r = nullptr; }
    };
    try {
        b->offsets.assign((size_t)B + 1, 0);
        for (int i = 0; i < B; ++i) b->offsets[(size_t)i + 1] = b->offsets[(size_t)i] + (int64_t)(reqs[(size_t)i]->n_bytes / 2);
        b->pcm.resize((size_t)std::max<int64_t>(b->offsets[(size_t)B], 1));
        // ragged layouts end to end: request i's features are the dense [128][features_len_i] block the reference hands its
        // encoder (src/triton/model.rs:126-141), its encoder output the [1024][encoded_len_i] block it gets back
        b->foff.assign((size_t)B + 1, 0);
        for (int i = 0; i < B; ++i) {
            std::memcpy(b->pcm.data() + b->offsets[(size_t)i], reqs[(size_t)i]->bytes, reqs[(size_t)i]->n_bytes);
            int64_t fl = 0;
            amira_features_len((int64_t)(reqs[(size_t)i]->n_bytes / 2), &fl);
            b->foff[(size_t)i + 1] = b->foff[(size_t)i] + (int64_t)AMIRA_N_MELS * fl;
        }
        b->features.resize((size_t)std::max<int64_t>(b->foff[(size_t)B], 1));
        b->flens.assign((size_t)B, 0);
        int32_t rc = amira_preprocess_pcm16_packed(p->ctx, b->pcm.data(), b->offsets.data(), B, b->features.data(), b->foff.data(), b->flens.data());
        if (rc) return fail_all(rc, amira_last_error(p->ctx));
        // encoder (out of scope, injected): per utterance, contract [1][128][features_len] -> [1][1024][encoded_len]
        if (!p->encoder) return fail_all(AMIRA_ERR_NOT_READY, "no encoder callback installed");
        b->elens.assign((size_t)B, 0);
        b->eoff.assign((size_t)B + 1, 0);
        b->enc.clear();
        for (int i = 0; i < B; ++i) {
            const float *e = nullptr;
            int64_t el = 0;
            if (p->encoder(p->encoder_user, b->features.data() + b->foff[(size_t)i], b->flens[(size_t)i], &e, &el) != 0 || el < 0 || (el > 0 && !e)) {
                finish_req(b, reqs[(size_t)i], AMIRA_ERR_UNKNOWN, "encoder callback failed");
                reqs[(size_t)i] = nullptr;  // this request is out; the rest of the batch goes on
                el = 0;
            }
            if (el > 0) b->enc.insert(b->enc.end(), e, e + (size_t)AMIRA_ENC_DIM * (size_t)el);  // the callback's buffer is only valid until its next call
            b->elens[(size_t)i] = el;
            b->eoff[(size_t)i + 1] = b->eoff[(size_t)i] + (int64_t)AMIRA_ENC_DIM * el;
        }
        int32_t row_cap = AMIRA_MAX_TOTAL_TOKENS;  // the decode entry writes max_total_tokens ids per stream
        amira_ctx_max_total_tokens(p->ctx, &row_cap);
        b->tokens.assign((size_t)B * (size_t)row_cap, 0);
        b->ntok.assign((size_t)B, 0);
        if (b->eoff[(size_t)B] > 0) {
            rc = amira_greedy_decode_packed(p->ctx, b->enc.data(), b->eoff.data(), B, b->elens.data(), nullptr, nullptr, b->tokens.data(),
                                            b->ntok.data(), nullptr);
            if (rc && rc != AMIRA_ERR_DECODE_STEP) return fail_all(rc, amira_last_error(p->ctx));
        }
        for (int i = 0; i < B; ++i) {
            BatchReq *r = reqs[(size_t)i];
            if (!r) continue;
            const int32_t n = b->ntok[(size_t)i];
            if (n < 0) {  // this stream's argmax left the embedding table ("Decode step failed", decoder_optimized.rs:148-152)
                finish_req(b, r, AMIRA_ERR_DECODE_STEP, "Decode step failed");
                reqs[(size_t)i] = nullptr;
                continue;
            }
            const int32_t *tk = b->tokens.data() + (size_t)i * (size_t)row_cap;
            r->out->audio_length_samples = (int64_t)(r->n_bytes / 2);
            r->out->features_length = b->flens[(size_t)i];
            r->out->encoded_length = b->elens[(size_t)i];
            r->out->n_tokens = n;
            if (r->tokens && r->tokens_cap > 0) std::memcpy(r->tokens, tk, sizeof(int32_t) * (size_t)std::min(n, r->tokens_cap));
            const std::string s = p->vocab.decode(tk, n);
            r->out->text_len = (int32_t)s.size();
            if (r->text && r->text_cap) {
                const size_t m = std::min(s.size(), r->text_cap - 1);
                std::memcpy(r->text, s.data(), m);
                r->text[m] = '\0';
            }
            reqs[(size_t)i] = nullptr;  // before the hand-back: the owner may free the request once it is done
            finish_req(b, r, AMIRA_OK, "");
        }
    } catch (const std::bad_alloc &) {
        fail_all(AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (...) {
        fail_all(AMIRA_ERR_UNKNOWN, "unexpected exception");
    }
}

void batcher_loop(amira_batcher *b) {
    std::unique_lock<std::mutex> lock(b->mu);
    for (;;) {
        b->cv_work.wait(lock, [&] { return b->stop || !b->queue.empty(); });
        if (b->stop && b->queue.empty()) return;
        // coalescing window: wait for more requests, up to max_wait_us after the first one or until the batch is full
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(b->max_wait_us);
        b->cv_work.wait_until(lock, deadline, [&] { return b->stop || (int)b->queue.size() >= b->max_batch; });
        std::vector<BatchReq *> reqs;
        while (!b->queue.empty() && (int)reqs.size() < b->max_batch) {
            reqs.push_back(b->queue.front());
            b->queue.pop_front();
        }
        lock.unlock();
        b->n_batches.fetch_add(1);
        run_batch(b, reqs);
        lock.lock();
    }
}

}  // namespace

extern "C" {

int32_t amira_batcher_create(amira_pipeline *p, int32_t max_batch, int32_t max_wait_us, amira_batcher **out) {
    if (!p || !out || max_batch <= 0 || max_wait_us < 0) return pfail(p, AMIRA_ERR_INVALID_VALUE, "amira_batcher_create: bad arguments");
    *out = nullptr;
    if (!p->ctx) return pfail(p, AMIRA_ERR_NO_DEVICE, "pipeline was created without a GPU context");
    try {
        amira_batcher *b = new amira_batcher();
        b->p = p;
        b->max_batch = max_batch;
        b->max_wait_us = max_wait_us;
        b->worker = std::thread(batcher_loop, b);
        *out = b;
    } catch (...) {
        return pfail(p, AMIRA_ERR_OUT_OF_MEMORY, "cannot start the batcher thread");
    }
    return AMIRA_OK;
}

int32_t amira_batcher_destroy(amira_batcher *b) {
    if (!b) return AMIRA_OK;
    {
        std::lock_guard<std::mutex> lock(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    if (b->worker.joinable()) b->worker.join();
    delete b;
    return AMIRA_OK;
}

int32_t amira_batcher_process_batch(amira_batcher *b, const uint8_t *audio_bytes, size_t n_bytes, amira_transcription *out,
                                    int32_t *tokens, int32_t tokens_cap, char *text, size_t text_cap) {
    if (!b || !out) return AMIRA_ERR_INVALID_VALUE;
    std::memset(out, 0, sizeof(*out));
    if (text && text_cap) text[0] = '\0';
    if (!audio_bytes || n_bytes == 0 || (n_bytes & 1))  // empty / odd-length requests take the single-request path and its rules
        return amira_pipeline_process_batch(b->p, audio_bytes, n_bytes, out, tokens, tokens_cap, text, text_cap);
    BatchReq r;
    r.bytes = audio_bytes; r.n_bytes = n_bytes; r.out = out; r.tokens = tokens; r.tokens_cap = tokens_cap; r.text = text; r.text_cap = text_cap;
    b->n_requests.fetch_add(1);
    std::unique_lock<std::mutex> lock(b->mu);
    if (b->stop) return AMIRA_ERR_NOT_READY;
    b->queue.push_back(&r);
    b->cv_work.notify_all();
    b->cv_done.wait(lock, [&] { return r.done; });
    lock.unlock();
    if (r.rc) {
        std::lock_guard<std::mutex> plock(b->p->mu);
        b->p->err = r.err;
    }
    return r.rc;
}

int32_t amira_batcher_stats(amira_batcher *b, int64_t *n_requests, int64_t *n_batches) {
    if (!b) return AMIRA_ERR_INVALID_VALUE;
    if (n_requests) *n_requests = b->n_requests.load();
    if (n_batches) *n_batches = b->n_batches.load();
    return AMIRA_OK;
}

}  // extern "C"
