// host_pipeline.cpp — C++ host side above the C ABI: the mirror of the reference's Rust `AsrPipeline` trait
// (src/asr/pipeline.rs:20-67) and of TritonAsrPipeline::process_audio_zero_copy (src/asr/pipeline.rs:269-380),
// with `preprocessor` and `decoder_joint` served by the GPU library instead of Triton.  Rust is not available in
// this build environment; rust/amira-b200-sys carries the equivalent crate as source.
// Also: Vocabulary (src/asr/types.rs:77-155) and the host-side utterance sharder for multi-GPU.
#include <algorithm>
#include <cstdio>
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

namespace {

bool is_space(unsigned char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\v' || ch == '\f'; }

// Vocabulary::load_from_file (src/asr/types.rs:87-108): "<token> <id>" per line; token = all fields but the last.
struct Vocabulary {
    std::unordered_map<int32_t, std::string> id_to_token;

    bool load(const std::string &path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) return false;
        std::string line;
        while (std::getline(f, line)) {
            std::vector<std::string> parts;
            size_t i = 0;
            while (i < line.size()) {
                while (i < line.size() && is_space((unsigned char)line[i])) ++i;
                size_t j = i;
                while (j < line.size() && !is_space((unsigned char)line[j])) ++j;
                if (j > i) parts.emplace_back(line.substr(i, j - i));
                i = j;
            }
            if (parts.size() < 2) continue;
            const std::string &id_s = parts.back();
            char *end = nullptr;
            errno = 0;
            const long long v = std::strtoll(id_s.c_str(), &end, 10);
            if (errno || end == id_s.c_str() || *end != '\0' || v < INT32_MIN || v > INT32_MAX) continue;
            if (id_s[0] == '+' && id_s.size() == 1) continue;
            std::string tok = parts[0];
            for (size_t k = 1; k + 1 < parts.size(); ++k) tok += " " + parts[k];
            id_to_token[(int32_t)v] = tok;  // later lines overwrite (HashMap::insert)
        }
        return true;
    }

    // Vocabulary::decode_tokens (src/asr/types.rs:111-135)
    std::string decode(const int32_t *ids, int32_t n) const {
        static const char kSp[] = "\xE2\x96\x81";  // U+2581
        std::string out;
        for (int32_t i = 0; i < n; ++i) {
            auto it = id_to_token.find(ids[i]);
            if (it == id_to_token.end()) continue;  // unknown ids are skipped silently (:115-116)
            const std::string &t = it->second;
            if (t.compare(0, 3, kSp) == 0) {
                out += ' ';
                out.append(t, 3, std::string::npos);
            } else {
                out += t;
            }
        }
        size_t a = 0, b = out.size();
        while (a < b && is_space((unsigned char)out[a])) ++a;
        while (b > a && is_space((unsigned char)out[b - 1])) --b;
        return out.substr(a, b - a);
    }
};

}  // namespace

struct amira_pipeline {
    amira_ctx *ctx = nullptr;
    amira_encoder_fn encoder = nullptr;
    void *encoder_user = nullptr;
    Vocabulary vocab;
    std::mutex mu;
    std::string err;
    std::vector<float> features, wave;
    std::vector<int32_t> tokens;
};

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
    p->tokens.assign(AMIRA_MAX_TOTAL_TOKENS * 8, 0);  // >= max_total_tokens of any sane config
    int32_t ntok = 0;
    if (enc_len > 0) {
        rc = amira_greedy_decode(p->ctx, enc, 1, (int32_t)enc_len, &enc_len, states_1, states_2, p->tokens.data(), &ntok, nullptr);
        if (rc) return pfail(p, rc, amira_last_error(p->ctx));
    }
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
