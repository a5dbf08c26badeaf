// api.cu — the C ABI of libamira_b200.so (include/amira_b200.h): context lifecycle, host<->device staging and
// argument validation around the kernels in frontend.cu / decoder.cu.  Conventions follow the reference's own
// FFI crate (src/cuda/mod.rs:54-62,371-412; src/cuda/cuda_helper.cu): status codes by value, opaque handle,
// no exception crosses the boundary.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

#include "common.h"

using namespace amira;

std::mutex &amira::upload_turn_mutex(int device) {
    static std::mutex m[64];
    return m[device & 63];
}
int amira::upload_fifo_mode() {
    static const int mode = getenv("AMIRA_H2D_FIFO") ? atoi(getenv("AMIRA_H2D_FIFO")) : 1;
    return mode;
}


struct amira_ctx : public amira::Ctx {};

namespace {

thread_local std::string g_create_error;

int32_t fail(Ctx *c, int32_t code, const std::string &msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
int32_t fail_cuda(Ctx *c, cudaError_t e, const char *what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear sticky-free errors
    return fail(c, e == cudaErrorMemoryAllocation ? AMIRA_ERR_OUT_OF_MEMORY : AMIRA_ERR_UNKNOWN, m);
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Input argument: device pointers are used in place, host pointers are copied into staging slot `slot`.
template <class T>
cudaError_t stage_in(Ctx *c, int slot, const T *p, size_t count, const T **dev) {
    if (!p || count == 0) {
        *dev = p && is_device_ptr(p) ? p : nullptr;
        if (p && !*dev) {  // zero-length host input: hand the kernels a valid dummy pointer
            cudaError_t e = c->stage[slot].reserve(16);
            if (e != cudaSuccess) return e;
            *dev = c->stage[slot].as<T>();
        }
        return cudaSuccess;
    }
    if (is_device_ptr(p)) {
        *dev = p;
        return cudaSuccess;
    }
    cudaError_t e = c->stage[slot].reserve(count * sizeof(T));
    if (e != cudaSuccess) return e;
    *dev = c->stage[slot].as<T>();
    return cudaMemcpyAsync(c->stage[slot].p, p, count * sizeof(T), cudaMemcpyHostToDevice, c->stream);
}

// Output argument: returns the device pointer kernels should write; finish_out copies back when `p` is host.
template <class T>
cudaError_t stage_out(Ctx *c, int slot, T *p, size_t count, T **dev, bool *is_host) {
    *is_host = false;
    if (!p) {
        *dev = nullptr;
        return cudaSuccess;
    }
    if (is_device_ptr(p)) {
        *dev = p;
        return cudaSuccess;
    }
    *is_host = true;
    cudaError_t e = c->stage[slot].reserve(count * sizeof(T) + 16);
    if (e != cudaSuccess) return e;
    *dev = c->stage[slot].as<T>();
    return cudaSuccess;
}
template <class T>
cudaError_t finish_out(Ctx *c, T *host, const T *dev, size_t count, bool is_host) {
    if (!is_host || !host || count == 0) return cudaSuccess;
    return cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream);
}

// A failing call must not leave copies in flight that still target the caller's host buffers or this context's pinned /
// staging buffers (the caller may free its buffers after the error, the next call reuses ours): every exit that is not the
// success path waits for the three streams first.
struct DrainOnError {
    Ctx *c;
    bool armed = true;
    explicit DrainOnError(Ctx *ctx) : c(ctx) {}
    ~DrainOnError() {
        if (!armed) return;
        if (c->h2d_stream) cudaStreamSynchronize(c->h2d_stream);
        if (c->stream) cudaStreamSynchronize(c->stream);
        if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
        cudaGetLastError();
    }
};

#define CK(call, what)                                      \
    do {                                                    \
        cudaError_t _e = (call);                            \
        if (_e != cudaSuccess) return fail_cuda(c, _e, what); \
    } while (0)

#define API_BEGIN(c)                                                                   \
    if (!(c)) return fail(nullptr, AMIRA_ERR_INVALID_VALUE, "null context");           \
    std::lock_guard<std::mutex> _lock((c)->mu);                                        \
    try {                                                                              \
        cudaError_t _se = cudaSetDevice((c)->device);                                  \
        if (_se != cudaSuccess) return fail_cuda((c), _se, "cudaSetDevice");

#define API_END(c)                                                                     \
    }                                                                                  \
    catch (const std::bad_alloc &) { return fail((c), AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed"); } \
    catch (const std::exception &ex) { return fail((c), AMIRA_ERR_UNKNOWN, ex.what()); } \
    catch (...) { return fail((c), AMIRA_ERR_UNKNOWN, "unknown exception"); }

}  // namespace

static void prof_collect(Ctx *c) {  // fold finished spans into the per-kernel totals (stream must be idle)
    for (auto &sp : c->prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.beg, sp.end) == cudaSuccess && sp.kernel >= 0 && sp.kernel < 8) {
            c->prof_ms[sp.kernel] += ms;
            c->prof_n[sp.kernel] += 1;
        }
        cudaEventDestroy(sp.beg);
        cudaEventDestroy(sp.end);
    }
    cudaGetLastError();
    c->prof_spans.clear();
}

extern "C" {

int32_t amira_device_count(int32_t *count) {
    if (!count) return AMIRA_ERR_INVALID_VALUE;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? AMIRA_OK : AMIRA_ERR_UNKNOWN;
    }
    *count = n;
    return AMIRA_OK;
}

// Recovery from a sticky device error.  The persistent kernels guard every spin loop with a cycle watchdog that traps instead of
// hanging the GPU; a trap (like any device-side fault) leaves the process's CUDA context on that device unusable: every later call
// of every amira_ctx on it fails with AMIRA_ERR_UNKNOWN.  The way back: destroy those contexts, call this, create new ones.
int32_t amira_device_reset(int32_t device_id) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device_id < 0 || device_id >= n) {
        cudaGetLastError();
        return AMIRA_ERR_INVALID_VALUE;
    }
    if (cudaSetDevice(device_id) != cudaSuccess) { cudaGetLastError(); return AMIRA_ERR_UNKNOWN; }
    const cudaError_t e = cudaDeviceReset();
    cudaGetLastError();
    return e == cudaSuccess ? AMIRA_OK : AMIRA_ERR_UNKNOWN;
}

int32_t amira_config_default(amira_config *cfg) {
    if (!cfg) return AMIRA_ERR_INVALID_VALUE;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->device_id = 0;
    cfg->max_symbols_per_step = AMIRA_MAX_SYMBOLS_PER_STEP;
    cfg->max_total_tokens = AMIRA_MAX_TOTAL_TOKENS;
    cfg->blank_id = AMIRA_BLANK_ID;
    cfg->joint_activation = 0;
    cfg->decode_engine = 0;
    cfg->max_streams = 1024;
    return AMIRA_OK;
}

int32_t amira_ctx_create(const amira_config *cfg, amira_ctx **out) {
    if (!out) return fail(nullptr, AMIRA_ERR_INVALID_VALUE, "null out pointer");
    *out = nullptr;
    amira_config dflt;
    amira_config_default(&dflt);
    if (!cfg) cfg = &dflt;
    if (cfg->max_symbols_per_step <= 0 || cfg->max_total_tokens <= 0 || cfg->blank_id < 0 || cfg->blank_id >= kV ||
        cfg->max_streams < 0 || cfg->joint_activation < 0 || cfg->joint_activation > 1 || (cfg->decode_engine != 0 && cfg->decode_engine != 1 && cfg->decode_engine != 4) ||
        cfg->decode_rule < 0 || cfg->decode_rule > (AMIRA_RULE_STATE_ON_NONBLANK | AMIRA_RULE_TDT_DURATIONS) || (cfg->decode_rule != 0 && cfg->decode_engine == 4))
        return fail(nullptr, AMIRA_ERR_INVALID_VALUE, "invalid amira_config");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, AMIRA_ERR_NO_DEVICE, "no CUDA device: libamira_b200 has no CPU path");
    }
    if (cfg->device_id < 0 || cfg->device_id >= n) return fail(nullptr, AMIRA_ERR_INVALID_VALUE, "device_id out of range");
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, cfg->device_id) != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, AMIRA_ERR_NO_DEVICE, "cannot query device");
    }
    if (prop.major != 10)
        return fail(nullptr, AMIRA_ERR_NO_DEVICE,
                    std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                        "; this library is built for sm_100a (B200) only");
    amira_ctx *c = nullptr;
    try {
        c = new amira_ctx();
    } catch (...) {
        return fail(nullptr, AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    c->cfg = *cfg;
    c->device = cfg->device_id;
    c->sm_count = prop.multiProcessorCount;
    auto bail = [&](cudaError_t e, const char *what) {
        int32_t rc = fail_cuda(nullptr, e, what);
        amira_ctx_destroy(c);
        return rc;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    c->stream = c->own_stream;
    if ((e = cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    for (cudaEvent_t &ev : c->ev_pool)  // blocking sync: a host thread throttled on one of these yields its core
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&c->ev_block, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) return bail(e, "event");
    try {  // host allocations below may throw: no exception crosses the C boundary, and the half-built context is released
        c->shared = std::make_shared<SharedDev>();
        c->shared->device = c->device;
        std::unique_ptr<FrontendTables> ht(new FrontendTables());
        build_frontend_tables(ht.get());
        e = cudaMalloc(&c->shared->tables_dev, sizeof(FrontendTables));
        if (e == cudaSuccess) e = cudaMemcpy(c->shared->tables_dev, ht.get(), sizeof(FrontendTables), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = frontend_upload_tables(ht.get());
        c->tables_dev = c->shared->tables_dev;
        if (cfg->max_streams > 0) c->slot_used.assign((size_t)cfg->max_streams, 0);
    } catch (...) {
        amira_ctx_destroy(c);
        return fail(nullptr, AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    if (e != cudaSuccess) return bail(e, "front-end tables");
    if (cfg->max_streams > 0) {
        const size_t bytes = sizeof(float) * (size_t)cfg->max_streams * 2 * kH;
        if ((e = cudaMalloc(&c->slot_s1, bytes)) != cudaSuccess) return bail(e, "stream slots");
        if ((e = cudaMalloc(&c->slot_s2, bytes)) != cudaSuccess) return bail(e, "stream slots");
        cudaMemset(c->slot_s1, 0, bytes);
        cudaMemset(c->slot_s2, 0, bytes);
    }
    *out = c;
    return AMIRA_OK;
}

int32_t amira_ctx_destroy(amira_ctx *c) {
    if (!c) return AMIRA_OK;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    decoder_release(c);  // this lane's workspace; tables and weights are freed with the last holder of `shared`
    if (c->h2d_stream) cudaStreamSynchronize(c->h2d_stream);
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    c->tables_dev = nullptr;
    c->w_blob = nullptr;
    c->shared.reset();
    if (c->slot_s1) cudaFree(c->slot_s1);
    if (c->slot_s2) cudaFree(c->slot_s2);
    for (int k = 0; k < Ctx::kMaxChunks; ++k) {
        c->fe_meta[k].release();
        c->fe_partials[k].release();
        c->fe_meta_pin[k].release();
    }
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (cudaEvent_t ev : c->ev_pool)
        if (ev) cudaEventDestroy(ev);
    if (c->ev_block) cudaEventDestroy(c->ev_block);
    prof_collect(c);
    for (auto &b : c->stage) b.release();
    for (auto &b : c->pin) b.release();
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    cudaGetLastError();
    delete c;
    return AMIRA_OK;
}

const char *amira_last_error(amira_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int32_t amira_ctx_set_stream(amira_ctx *c, void *cuda_stream) {
    API_BEGIN(c)
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_ctx_synchronize(amira_ctx *c) {
    API_BEGIN(c)
    CK(cudaStreamSynchronize(c->stream), "synchronize");
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_ctx_launch_count(amira_ctx *c, int64_t *count) {
    if (!c || !count) return AMIRA_ERR_INVALID_VALUE;
    *count = c->launches;
    return AMIRA_OK;
}

int32_t amira_ctx_max_total_tokens(amira_ctx *c, int32_t *value) {
    if (!c || !value) return AMIRA_ERR_INVALID_VALUE;
    *value = c->cfg.max_total_tokens;
    return AMIRA_OK;
}

int32_t amira_ctx_profile(amira_ctx *c, int32_t enable) {
    API_BEGIN(c)
    CK(cudaStreamSynchronize(c->stream), "profile sync");
    prof_collect(c);
    for (int i = 0; i < 8; ++i) { c->prof_ms[i] = 0.0; c->prof_n[i] = 0; }
    c->profiling = enable != 0;
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_ctx_kernel_ms(amira_ctx *c, int32_t kernel, double *total_ms, int64_t *launches) {
    API_BEGIN(c)
    if (kernel < 0 || kernel >= PK_COUNT || !total_ms || !launches) return fail(c, AMIRA_ERR_INVALID_VALUE, "bad kernel id");
    CK(cudaStreamSynchronize(c->stream), "profile sync");
    prof_collect(c);
    *total_ms = c->prof_ms[kernel];
    *launches = c->prof_n[kernel];
    return AMIRA_OK;
    API_END(c)
}

// ---------------------------------------------------------------------------------------------- weights
int32_t amira_weights_random_init(float *blob, size_t n_params, uint64_t seed, float blank_bias) {
    if (!blob || n_params != (size_t)AMIRA_N_PARAMS) return AMIRA_ERR_INVALID_VALUE;
    weights_random_init(blob, seed, blank_bias);
    return AMIRA_OK;
}

int32_t amira_ctx_load_weights(amira_ctx *c, const float *blob, size_t n_params) {
    API_BEGIN(c)
    if (!blob || n_params != (size_t)AMIRA_N_PARAMS || blob_layout().total != (size_t)AMIRA_N_PARAMS)
        return fail(c, AMIRA_ERR_INVALID_VALUE, "weight blob must hold exactly AMIRA_N_PARAMS fp32 values");
    // the weights live in `shared`: every lane forked from this context sees the new ones at its next call (the caller keeps the
    // other lanes idle during a reload, as it would for any model swap)
    std::lock_guard<std::mutex> wlock(c->shared->mu);
    if (!c->shared->w_blob) CK(cudaMalloc(&c->shared->w_blob, sizeof(float) * n_params), "weights alloc");
    c->w_blob = c->shared->w_blob;
    CK(cudaMemcpyAsync(c->w_blob, blob, sizeof(float) * n_params,
                       is_device_ptr(blob) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream),
       "weights copy");
    CK(decoder_prepare_weights(c), "derived weight tables");
    CK(cudaStreamSynchronize(c->stream), "weights sync");
    c->has_weights = true;
    return AMIRA_OK;
    API_END(c)
}

// More submission lanes on the same GPU: the new context has its own streams, staging buffers and decode workspace and shares the
// parent's weights and tables (read-only).  Calls on different lanes run concurrently: the upload of one batch overlaps the
// kernels of another without a second copy of the model.  Stream slots stay with the parent (max_streams = 0 in the fork).
int32_t amira_ctx_fork(amira_ctx *parent, amira_ctx **out) {
    if (!out) return fail(parent, AMIRA_ERR_INVALID_VALUE, "null out pointer");
    *out = nullptr;
    if (!parent) return fail(nullptr, AMIRA_ERR_INVALID_VALUE, "null context");
    amira_config cfg;
    std::shared_ptr<SharedDev> sh;
    {
        std::lock_guard<std::mutex> lock(parent->mu);
        cfg = parent->cfg;
        sh = parent->shared;
    }
    cfg.max_streams = 0;
    amira_ctx *c = nullptr;
    const int32_t rc = amira_ctx_create(&cfg, &c);
    if (rc) { parent->err = g_create_error; return rc; }
    try {
        std::lock_guard<std::mutex> lock(c->mu);
        c->shared = sh;  // drops the fork's own (tables only) shared block
        c->tables_dev = sh->tables_dev;
        std::lock_guard<std::mutex> wlock(sh->mu);
        decoder_adopt_shared(c);
    } catch (...) {
        amira_ctx_destroy(c);
        return fail(parent, AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    *out = c;
    return AMIRA_OK;
}

int32_t amira_ctx_load_weights_file(amira_ctx *c, const char *path) {
    if (!c) return AMIRA_ERR_INVALID_VALUE;
    if (!path) return fail(c, AMIRA_ERR_INVALID_VALUE, "null path");
    std::vector<float> blob;
    try {
        blob.resize(AMIRA_N_PARAMS);
    } catch (...) {
        return fail(c, AMIRA_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(c, AMIRA_ERR_IO, std::string("cannot open ") + path);
    const size_t got = std::fread(blob.data(), sizeof(float), blob.size(), f);
    const bool extra = std::fgetc(f) != EOF;
    std::fclose(f);
    if (got != blob.size() || extra) return fail(c, AMIRA_ERR_IO, "weight file must hold exactly AMIRA_N_PARAMS fp32 values");
    return amira_ctx_load_weights(c, blob.data(), blob.size());
}

// End-of-call wait.  Batch-sized calls take tens of milliseconds of GPU time: the host thread sleeps on a blocking-sync event
// (a server runs many such threads on few cores; spinning ones starve the threads that feed the copy engines).  Small calls
// (single requests, streaming ticks) keep the spinning wait and its microsecond wake-up.
static cudaError_t wait_stream(Ctx *c, cudaStream_t s, bool blocking) {
    static const int spin_env = getenv("AMIRA_SPIN_WAIT") ? atoi(getenv("AMIRA_SPIN_WAIT")) : 0;  // A/B timing of the wake-up cost
    if (!blocking || spin_env) return cudaStreamSynchronize(s);
    cudaError_t e = cudaEventRecord(c->ev_block, s);
    return e != cudaSuccess ? e : cudaEventSynchronize(c->ev_block);
}

// ---------------------------------------------------------------------------------------------- front end
int32_t amira_features_len(int64_t n_samples, int64_t *features_len) {
    if (!features_len) return AMIRA_ERR_INVALID_VALUE;
    *features_len = n_samples <= 0 ? 0 : n_samples / kHop + 1;
    return AMIRA_OK;
}

static int32_t preprocess_common(amira_ctx *c, const void *wave, bool pcm16, const int64_t *starts, const int64_t *lens,
                                 int64_t total_elems, int32_t B, float *features, int64_t t_stride,
                                 int64_t *features_lens, const int64_t *feat_offsets = nullptr, bool normalize = true) {
    // feat_offsets != nullptr: packed output — utterance b is a [128][features_len_b] block at features + feat_offsets[b]
    int64_t max_len = 0;
    for (int b = 0; b < B; ++b) {
        if (lens[b] < 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "negative waveform length");
        const int64_t L = lens[b] <= 0 ? 0 : lens[b] / kHop + 1;
        if (features_lens) features_lens[b] = L;
        max_len = L > max_len ? L : max_len;
    }
    if (feat_offsets) {
        t_stride = 0;
        if (feat_offsets[0] < 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "feat_offsets must be non-negative");
        for (int b = 0; b < B; ++b) {
            const int64_t L = lens[b] <= 0 ? 0 : lens[b] / kHop + 1;
            if (feat_offsets[b + 1] - feat_offsets[b] < (int64_t)kMel * L)
                return fail(c, AMIRA_ERR_INVALID_VALUE, "feat_offsets: block smaller than 128 x features_len");
        }
    } else if (t_stride < max_len || t_stride <= 0) {
        return fail(c, AMIRA_ERR_INVALID_VALUE, "t_stride smaller than the longest features_len");
    }
    DrainOnError drain(c);
    const size_t esz = pcm16 ? sizeof(int16_t) : sizeof(float);
    const size_t feat_count = feat_offsets ? (size_t)feat_offsets[B] : (size_t)B * kMel * (size_t)t_stride;
    auto feat_elem = [&](int b) -> size_t { return feat_offsets ? (size_t)feat_offsets[b] : (size_t)b * kMel * (size_t)t_stride; };
    const bool wave_host = wave && total_elems > 0 && !is_device_ptr(wave), feat_host = !is_device_ptr(features);
    const uint8_t *wave_dev = static_cast<const uint8_t *>(wave);
    float *feat_dev = features;
    if (wave_host || !wave || total_elems == 0) {
        CK(c->stage[0].reserve((size_t)total_elems * esz + 16), "waveform staging");
        wave_dev = c->stage[0].as<uint8_t>();
    }
    if (feat_host) {
        CK(c->stage[1].reserve(feat_count * sizeof(float) + 16), "features staging");
        feat_dev = c->stage[1].as<float>();
    }
    // Host buffers: utterances are independent, so the call is pipelined in chunks of utterances — H2D of chunk k+1, the
    // kernels of chunk k and the D2H of chunk k-1 run concurrently on three streams.  Device buffers: one chunk, no copies.
    // (a chunk should carry at least a couple of megabytes; measured on a 1024-stream 160 ms tick — 5 MB in, 9-17 MB out — the
    // chunked pipeline still beats a single chunk by 0.2 ms, because the feature download dominates and overlaps the kernels)
    const size_t moved = (wave_host ? (size_t)total_elems * esz : 0) + (feat_host ? feat_count * sizeof(float) : 0);
    const int by_bytes = (int)std::min<size_t>((size_t)Ctx::kMaxChunks, moved / ((size_t)2 << 20));
    const int n_chunks = (wave_host || feat_host) ? std::max(1, std::min(by_bytes, B / 32)) : 1;
    const bool dbg = std::getenv("AMIRA_DEBUG_TIMELINE") != nullptr;
    cudaEvent_t dbg_ev[1 + 3 * Ctx::kMaxChunks] = {};
    if (dbg) {
        for (auto &ev : dbg_ev) cudaEventCreate(&ev);
        cudaEventRecord(dbg_ev[0], c->stream);
    }
    auto chunk_elem = [&](int b) -> int64_t { return b >= B ? total_elems : starts[b]; };
    for (int k = 0; k < n_chunks && n_chunks > 1; ++k) {  // every chunk's metadata first (see launch_frontend)
        const int b0 = (int)((int64_t)B * k / n_chunks), b1 = (int)((int64_t)B * (k + 1) / n_chunks);
        if (b1 > b0)
            CK(launch_frontend(c, wave_dev, pcm16, starts + b0, lens + b0, b1 - b0, feat_offsets ? feat_dev : feat_dev + feat_elem(b0), t_stride,
                               k, 1, feat_offsets ? feat_offsets + b0 : nullptr, normalize),
               "front-end metadata");
    }
    for (int k = 0; k < n_chunks; ++k) {
        const int b0 = (int)((int64_t)B * k / n_chunks), b1 = (int)((int64_t)B * (k + 1) / n_chunks);
        if (b1 <= b0) continue;
        if (wave_host) {
            const int64_t e0 = chunk_elem(b0), e1 = chunk_elem(b1);
            if (e1 > e0)
                CK(cudaMemcpyAsync(c->stage[0].as<uint8_t>() + (size_t)e0 * esz, static_cast<const uint8_t *>(wave) + (size_t)e0 * esz,
                                   (size_t)(e1 - e0) * esz, cudaMemcpyHostToDevice, c->h2d_stream), "waveform H2D");
            CK(cudaEventRecord(c->ev_pool[2 * k], c->h2d_stream), "event");
            CK(cudaStreamWaitEvent(c->stream, c->ev_pool[2 * k], 0), "event wait");
            if (dbg) cudaEventRecord(dbg_ev[1 + 3 * k], c->h2d_stream);
        }
        CK(launch_frontend(c, wave_dev, pcm16, starts + b0, lens + b0, b1 - b0, feat_offsets ? feat_dev : feat_dev + feat_elem(b0), t_stride, k,
                           n_chunks > 1 ? 2 : 0, feat_offsets ? feat_offsets + b0 : nullptr, normalize),
           "front-end launch");
        if (feat_host) {
            CK(cudaEventRecord(c->ev_pool[2 * k + 1], c->stream), "event");
            CK(cudaStreamWaitEvent(c->d2h_stream, c->ev_pool[2 * k + 1], 0), "event wait");
            const size_t o = feat_elem(b0), n = (b1 >= B ? feat_count : feat_elem(b1)) - o;
            if (n > 0) CK(cudaMemcpyAsync(features + o, feat_dev + o, n * sizeof(float), cudaMemcpyDeviceToHost, c->d2h_stream), "features D2H");
            if (dbg) { cudaEventRecord(dbg_ev[2 + 3 * k], c->stream); cudaEventRecord(dbg_ev[3 + 3 * k], c->d2h_stream); }
        }
    }
    // sleep only through calls that keep the GPU busy for milliseconds (>= 48 M samples = 50 minutes of audio): the wake-up of a
    // blocking-sync event costs ~0.27 ms, a third of a 64 x 30 s call (0.88 -> 0.62 ms) and a tenth of a 1024-stream 160 ms tick
    const bool sleep_wait = (int64_t)total_elems >= ((int64_t)48 << 20);
    if (feat_host) CK(wait_stream(c, c->d2h_stream, sleep_wait), "features D2H sync");
    CK(wait_stream(c, c->stream, sleep_wait), "front-end sync");
    if (dbg) {
        for (int k = 0; k < n_chunks && wave_host && feat_host; ++k) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, dbg_ev[0], dbg_ev[1 + 3 * k]);
            cudaEventElapsedTime(&b, dbg_ev[0], dbg_ev[2 + 3 * k]);
            cudaEventElapsedTime(&d, dbg_ev[0], dbg_ev[3 + 3 * k]);
            std::fprintf(stderr, "[fe timeline] chunk %d: H2D done %.2f ms, kernels done %.2f ms, D2H done %.2f ms\n", k, a, b, d);
        }
        for (auto &ev : dbg_ev) cudaEventDestroy(ev);
    }
    drain.armed = false;
    return AMIRA_OK;
}

int32_t amira_preprocess_pcm16(amira_ctx *c, const int16_t *pcm, const int64_t *offsets, int32_t B, float *features,
                               int64_t t_stride, int64_t *features_lens) {
    API_BEGIN(c)
    if (B < 0 || !offsets || !features || (B > 0 && !pcm && offsets[B] > 0))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_preprocess_pcm16: bad arguments");
    if (B == 0) return AMIRA_OK;
    std::vector<int64_t> lens((size_t)B);
    for (int b = 0; b < B; ++b) {
        lens[b] = offsets[b + 1] - offsets[b];
        if (lens[b] < 0 || offsets[b] < 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "offsets must be non-decreasing");
    }
    return preprocess_common(c, pcm, true, offsets, lens.data(), offsets[B], B, features, t_stride, features_lens);
    API_END(c)
}

int32_t amira_preprocess_pcm16_packed(amira_ctx *c, const int16_t *pcm, const int64_t *offsets, int32_t B, float *features,
                                      const int64_t *feat_offsets, int64_t *features_lens) {
    API_BEGIN(c)
    if (B < 0 || !offsets || !features || !feat_offsets || (B > 0 && !pcm && offsets[B] > 0))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_preprocess_pcm16_packed: bad arguments");
    if (B == 0) return AMIRA_OK;
    std::vector<int64_t> lens((size_t)B);
    for (int b = 0; b < B; ++b) {
        lens[b] = offsets[b + 1] - offsets[b];
        if (lens[b] < 0 || offsets[b] < 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "offsets must be non-decreasing");
    }
    return preprocess_common(c, pcm, true, offsets, lens.data(), offsets[B], B, features, 0, features_lens, feat_offsets);
    API_END(c)
}

// Un-normalised log-mel, ragged layout: log(mel + 2^-24) of every frame, the tensor the preprocessor holds before its per-feature
// normalisation.  The incremental streaming path normalises it with running statistics (host_stream.cpp).
int32_t amira_logmel_pcm16_packed(amira_ctx *c, const int16_t *pcm, const int64_t *offsets, int32_t B, float *features,
                                  const int64_t *feat_offsets, int64_t *features_lens) {
    API_BEGIN(c)
    if (B < 0 || !offsets || !features || !feat_offsets || (B > 0 && !pcm && offsets[B] > 0))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_logmel_pcm16_packed: bad arguments");
    if (B == 0) return AMIRA_OK;
    std::vector<int64_t> lens((size_t)B);
    for (int b = 0; b < B; ++b) {
        lens[b] = offsets[b + 1] - offsets[b];
        if (lens[b] < 0 || offsets[b] < 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "offsets must be non-decreasing");
    }
    return preprocess_common(c, pcm, true, offsets, lens.data(), offsets[B], B, features, 0, features_lens, feat_offsets, false);
    API_END(c)
}

int32_t amira_preprocess_f32(amira_ctx *c, const float *waveforms, int64_t n_stride, const int64_t *waveforms_lens,
                             int32_t B, float *features, int64_t t_stride, int64_t *features_lens) {
    API_BEGIN(c)
    if (B < 0 || !waveforms_lens || !features || n_stride < 0 || (B > 0 && n_stride > 0 && !waveforms))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_preprocess_f32: bad arguments");
    if (B == 0) return AMIRA_OK;
    std::vector<int64_t> starts((size_t)B);
    for (int b = 0; b < B; ++b) {
        starts[b] = (int64_t)b * n_stride;
        if (waveforms_lens[b] > n_stride) return fail(c, AMIRA_ERR_INVALID_VALUE, "waveforms_lens exceeds n_stride");
    }
    return preprocess_common(c, waveforms, false, starts.data(), waveforms_lens, (int64_t)B * n_stride, B, features,
                             t_stride, features_lens);
    API_END(c)
}

int32_t amira_preprocess_f32_packed(amira_ctx *c, const float *waveforms, const int64_t *wave_offsets, int32_t B, float *features,
                                    const int64_t *feat_offsets, int64_t *features_lens) {
    API_BEGIN(c)
    if (B < 0 || !wave_offsets || !features || !feat_offsets || (B > 0 && !waveforms && wave_offsets[B] > 0))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_preprocess_f32_packed: bad arguments");
    if (B == 0) return AMIRA_OK;
    std::vector<int64_t> lens((size_t)B);
    for (int b = 0; b < B; ++b) {
        lens[b] = wave_offsets[b + 1] - wave_offsets[b];
        if (lens[b] < 0 || wave_offsets[b] < 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "wave_offsets must be non-decreasing");
    }
    return preprocess_common(c, waveforms, false, wave_offsets, lens.data(), wave_offsets[B], B, features, 0, features_lens, feat_offsets);
    API_END(c)
}

int32_t amira_bytes_to_f32(amira_ctx *c, const uint8_t *bytes, size_t n_bytes, int32_t drop_odd, float *out,
                           size_t *n_out) {
    API_BEGIN(c)
    const size_t n = n_bytes / 2 + ((n_bytes & 1) && !drop_odd ? 1 : 0);
    if (n_out) *n_out = n;
    if (n_bytes == 0) return AMIRA_OK;
    if (!bytes || !out) return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_bytes_to_f32: null buffer");
    const uint8_t *in_dev = nullptr;
    CK(stage_in<uint8_t>(c, 0, bytes, n_bytes, &in_dev), "bytes H2D");
    float *out_dev = nullptr;
    bool out_host = false;
    CK(stage_out<float>(c, 1, out, n + 4, &out_dev, &out_host), "samples staging");
    CK(launch_bytes_to_f32(c, in_dev, n_bytes, drop_odd != 0, out_dev), "bytes_to_f32 launch");
    CK(finish_out<float>(c, out, out_dev, n, out_host), "samples D2H");
    CK(cudaStreamSynchronize(c->stream), "bytes_to_f32 sync");
    return AMIRA_OK;
    API_END(c)
}

// ---------------------------------------------------------------------------------------------- decoder_joint
int32_t amira_decoder_joint(amira_ctx *c, const float *encoder_outputs, int32_t B, int32_t T, const int32_t *targets,
                            int32_t U, const int32_t *target_length, const float *input_states_1,
                            const float *input_states_2, float *outputs, int32_t *prednet_lengths,
                            float *output_states_1, float *output_states_2) {
    API_BEGIN(c)
    if (c->shared && c->shared->version != c->weights_version) decoder_adopt_shared(c);
    if (!c->has_weights) return fail(c, AMIRA_ERR_NO_WEIGHTS, "amira_decoder_joint: no weights loaded");
    if (B <= 0 || T <= 0 || U <= 0 || !encoder_outputs || !targets || !outputs)
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_decoder_joint: bad arguments");
    const size_t n_state = (size_t)2 * B * kH;
    const float *enc_dev, *s1_dev, *s2_dev;
    const int32_t *tg_dev, *tl_dev;
    CK(stage_in<float>(c, 0, encoder_outputs, (size_t)B * kEnc * T, &enc_dev), "encoder_outputs H2D");
    CK(stage_in<int32_t>(c, 1, targets, (size_t)B * U, &tg_dev), "targets H2D");
    CK(stage_in<int32_t>(c, 2, target_length, (size_t)B, &tl_dev), "target_length H2D");
    CK(stage_in<float>(c, 3, input_states_1, n_state, &s1_dev), "input_states_1 H2D");
    CK(stage_in<float>(c, 4, input_states_2, n_state, &s2_dev), "input_states_2 H2D");
    float *out_dev, *o1_dev, *o2_dev;
    int32_t *pl_dev;
    bool out_h, o1_h, o2_h, pl_h;
    const size_t n_out = (size_t)B * U * T * kV;
    CK(stage_out<float>(c, 5, outputs, n_out, &out_dev, &out_h), "outputs staging");
    CK(stage_out<float>(c, 6, output_states_1, n_state, &o1_dev, &o1_h), "output_states_1 staging");
    CK(stage_out<float>(c, 7, output_states_2, n_state, &o2_dev, &o2_h), "output_states_2 staging");
    CK(stage_out<int32_t>(c, 8, prednet_lengths, (size_t)B, &pl_dev, &pl_h), "prednet_lengths staging");
    CK(c->stage[9].reserve(16), "flag alloc");
    int32_t *flag_dev = c->stage[9].as<int32_t>();
    CK(cudaMemsetAsync(flag_dev, 0, sizeof(int32_t), c->stream), "flag reset");
    CK(launch_decoder_joint(c, enc_dev, B, T, tg_dev, U, tl_dev, s1_dev, s2_dev, out_dev, pl_dev, o1_dev, o2_dev, flag_dev),
       "decoder_joint launch");
    CK(finish_out<float>(c, outputs, out_dev, n_out, out_h), "outputs D2H");
    CK(finish_out<float>(c, output_states_1, o1_dev, n_state, o1_h), "output_states_1 D2H");
    CK(finish_out<float>(c, output_states_2, o2_dev, n_state, o2_h), "output_states_2 D2H");
    CK(finish_out<int32_t>(c, prednet_lengths, pl_dev, (size_t)B, pl_h), "prednet_lengths D2H");
    int32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, flag_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream), "flag D2H");
    CK(cudaStreamSynchronize(c->stream), "decoder_joint sync");
    // an out-of-range target is an ONNX Gather failure in the reference => "Decode step failed"
    if (flag) return fail(c, AMIRA_ERR_DECODE_STEP, "Decode step failed: target id outside the embedding table");
    return AMIRA_OK;
    API_END(c)
}

// ---------------------------------------------------------------------------------------------- greedy decode
static int32_t greedy_common(amira_ctx *c, const float *encoder_outputs, int32_t B, int32_t T,
                             const int64_t *encoded_lengths, const int32_t *slots_host, float *states_1, float *states_2,
                             int32_t *tokens, int32_t *n_tokens, int32_t *n_steps, const int64_t *enc_offsets = nullptr,
                             int32_t *last_tokens = nullptr) {
    // enc_offsets != nullptr: packed encoder outputs — stream b is a [1024][encoded_lengths[b]] block at encoder_outputs + enc_offsets[b]
    if (c->shared && c->shared->version != c->weights_version) decoder_adopt_shared(c);  // a sibling lane loaded weights
    if (!c->has_weights) return fail(c, AMIRA_ERR_NO_WEIGHTS, "greedy decode: no weights loaded");
    if (B < 0 || T < 0 || !tokens || !n_tokens || (B > 0 && T > 0 && !encoder_outputs))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "greedy decode: bad arguments");
    if (B == 0) return AMIRA_OK;
    const int cap = c->cfg.max_total_tokens;
    // lengths (+ slots) -> one pinned block
    CK(c->pin[0].reserve(sizeof(int32_t) * 2 * (size_t)B), "lens pin");
    int32_t *h_lens = c->pin[0].as<int32_t>();
    int32_t *h_slots = h_lens + B;
    auto unmark = [&](int upto) {  // slot_used == 2 marks "seen in this tick" (duplicate detection in O(B))
        if (slots_host)
            for (int q = 0; q < upto; ++q) c->slot_used[(size_t)h_slots[q]] = 1;
    };
    for (int b = 0; b < B; ++b) {
        const int64_t L = encoded_lengths ? encoded_lengths[b] : T;
        if (L < 0 || L > T) { unmark(b); return fail(c, AMIRA_ERR_INVALID_VALUE, "encoded_lengths out of range"); }
        h_lens[b] = (int32_t)L;
        if (slots_host) {
            const int32_t s = slots_host[b];
            if (s < 0 || s >= c->cfg.max_streams || !c->slot_used[(size_t)s]) {
                unmark(b);
                return fail(c, AMIRA_ERR_INVALID_VALUE, "stream slot not open");
            }
            if (c->slot_used[(size_t)s] == 2) {
                unmark(b);
                return fail(c, AMIRA_ERR_INVALID_VALUE, "duplicate stream slot in one tick");
            }
            c->slot_used[(size_t)s] = 2;
            h_slots[b] = s;
        }
    }
    unmark(B);
    if (enc_offsets) {
        if (c->cfg.decode_engine == 1 || c->cfg.decode_rule != 0)
            return fail(c, AMIRA_ERR_INVALID_VALUE, "packed encoder outputs need a tcgen05 decode engine (and the reference's decode rule)");
        if (enc_offsets[0] < 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "enc_offsets must be non-negative");
        for (int b = 0; b < B; ++b)
            if (enc_offsets[b + 1] - enc_offsets[b] < (int64_t)kEnc * h_lens[b])
                return fail(c, AMIRA_ERR_INVALID_VALUE, "enc_offsets: block smaller than 1024 x encoded_length");
    }
    DrainOnError drain(c);
    CK(c->stage[9].reserve(sizeof(int32_t) * 2 * (size_t)B), "lens dev");
    CK(cudaMemcpyAsync(c->stage[9].p, h_lens, sizeof(int32_t) * 2 * (size_t)B, cudaMemcpyHostToDevice, c->stream), "lens H2D");
    const int32_t *lens_dev = c->stage[9].as<int32_t>();
    const int32_t *slots_dev = slots_host ? lens_dev + B : nullptr;

    // encoder outputs in host memory are uploaded by the launcher, chunk by chunk, overlapped with their projection
    const float *enc_dev = encoder_outputs, *enc_host = nullptr;
    if (encoder_outputs && T > 0 && !is_device_ptr(encoder_outputs)) {
        CK(c->stage[0].reserve(sizeof(float) * (enc_offsets ? (size_t)enc_offsets[B] : (size_t)B * kEnc * T) + 16), "encoder_outputs staging");
        enc_dev = c->stage[0].as<float>();
        enc_host = encoder_outputs;
    }
    const size_t n_state = (size_t)2 * B * kH;
    float *s1_dev = nullptr, *s2_dev = nullptr;
    bool s1_h = false, s2_h = false;
    if (!slots_host && states_1 && states_2) {
        CK(stage_out<float>(c, 1, states_1, n_state, &s1_dev, &s1_h), "states_1 staging");
        CK(stage_out<float>(c, 2, states_2, n_state, &s2_dev, &s2_h), "states_2 staging");
        if (s1_h) CK(cudaMemcpyAsync(s1_dev, states_1, sizeof(float) * n_state, cudaMemcpyHostToDevice, c->stream), "states_1 H2D");
        if (s2_h) CK(cudaMemcpyAsync(s2_dev, states_2, sizeof(float) * n_state, cudaMemcpyHostToDevice, c->stream), "states_2 H2D");
    } else if (!slots_host && (states_1 || states_2)) {
        return fail(c, AMIRA_ERR_INVALID_VALUE, "states_1 and states_2 must both be given or both be null");
    }
    // resume form: the token each stream emitted last (the LSTM input of its first step) in, the one it emitted last now out
    int32_t *last_dev = nullptr;
    bool last_h = false;
    if (last_tokens) {
        if (c->cfg.decode_engine == 1 || c->cfg.decode_rule != 0)
            return fail(c, AMIRA_ERR_INVALID_VALUE, "the resume form needs the tcgen05 decode engine (and the reference's decode rule)");
        CK(stage_out<int32_t>(c, 6, last_tokens, (size_t)B, &last_dev, &last_h), "last_tokens staging");
        if (last_h) {
            for (int b = 0; b < B; ++b)
                if (last_tokens[b] < 0 || last_tokens[b] >= kEmbRows) return fail(c, AMIRA_ERR_INVALID_VALUE, "last_tokens outside the embedding table");
            CK(cudaMemcpyAsync(last_dev, last_tokens, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, c->stream), "last_tokens H2D");
        }
    }
    int32_t *tok_dev, *nt_dev, *ns_dev;
    bool tok_h, nt_h, ns_h;
    CK(stage_out<int32_t>(c, 3, tokens, (size_t)B * cap, &tok_dev, &tok_h), "tokens staging");
    CK(stage_out<int32_t>(c, 4, n_tokens, (size_t)B, &nt_dev, &nt_h), "n_tokens staging");
    CK(stage_out<int32_t>(c, 5, n_steps, (size_t)B, &ns_dev, &ns_h), "n_steps staging");
    CK(launch_greedy_decode(c, enc_dev, enc_host, B, T, lens_dev, h_lens, slots_dev, s1_dev, s2_dev, tok_dev, nt_dev, ns_dev, enc_offsets, last_dev),
       "greedy decode launch");
    CK(finish_out<int32_t>(c, last_tokens, last_dev, (size_t)B, last_h), "last_tokens D2H");
    CK(finish_out<int32_t>(c, tokens, tok_dev, (size_t)B * cap, tok_h), "tokens D2H");
    CK(finish_out<int32_t>(c, n_tokens, nt_dev, (size_t)B, nt_h), "n_tokens D2H");
    CK(finish_out<int32_t>(c, n_steps, ns_dev, (size_t)B, ns_h), "n_steps D2H");
    CK(finish_out<float>(c, states_1, s1_dev, n_state, s1_h), "states_1 D2H");
    CK(finish_out<float>(c, states_2, s2_dev, n_state, s2_h), "states_2 D2H");
    int32_t n_failed = 0;
    CK(cudaMemcpyAsync(&n_failed, decoder_fail_count_dev(c), sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream), "status D2H");
    CK(wait_stream(c, c->stream, B >= 64 && !slots_host), "greedy decode sync");
    drain.armed = false;
    // streams whose argmax left the embedding table and needed another step: the reference's next decoder_joint
    // call fails ("Decode step failed", src/asr/decoder_optimized.rs:148-152); their n_tokens is -1
    if (n_failed > 0) return fail(c, AMIRA_ERR_DECODE_STEP, "Decode step failed for " + std::to_string(n_failed) + " stream(s)");
    return AMIRA_OK;
}

int32_t amira_greedy_decode(amira_ctx *c, const float *encoder_outputs, int32_t B, int32_t T,
                            const int64_t *encoded_lengths, float *states_1, float *states_2, int32_t *tokens,
                            int32_t *n_tokens, int32_t *n_steps) {
    API_BEGIN(c)
    return greedy_common(c, encoder_outputs, B, T, encoded_lengths, nullptr, states_1, states_2, tokens, n_tokens, n_steps);
    API_END(c)
}

int32_t amira_greedy_decode_packed(amira_ctx *c, const float *encoder_outputs, const int64_t *enc_offsets, int32_t B,
                                   const int64_t *encoded_lengths, float *states_1, float *states_2, int32_t *tokens,
                                   int32_t *n_tokens, int32_t *n_steps) {
    API_BEGIN(c)
    if (B < 0 || (B > 0 && (!enc_offsets || !encoded_lengths)))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_greedy_decode_packed: bad arguments");
    int64_t T = 0;
    for (int b = 0; b < B; ++b) {
        if (encoded_lengths[b] < 0 || encoded_lengths[b] > 0x7fffffff) return fail(c, AMIRA_ERR_INVALID_VALUE, "encoded_lengths out of range");
        T = encoded_lengths[b] > T ? encoded_lengths[b] : T;
    }
    return greedy_common(c, encoder_outputs, B, (int32_t)T, encoded_lengths, nullptr, states_1, states_2, tokens, n_tokens, n_steps,
                         enc_offsets);
    API_END(c)
}

// The loop resumed where an earlier call left it: besides the LSTM state, the token emitted last is carried (last_tokens [B],
// in/out, host or device; AMIRA_BLANK_ID for a fresh stream), so chunk-by-chunk decoding of one stream emits exactly the
// tokens of one call over the concatenated frames.  The reference carries the state only (src/asr/decoder_optimized.rs:78: every
// call starts from blank); this is the "carry state + last token" of a truly incremental decoder.
int32_t amira_greedy_decode_resume(amira_ctx *c, const float *encoder_outputs, const int64_t *enc_offsets, int32_t B,
                                   const int64_t *encoded_lengths, float *states_1, float *states_2, int32_t *last_tokens,
                                   int32_t *tokens, int32_t *n_tokens, int32_t *n_steps) {
    API_BEGIN(c)
    if (B < 0 || (B > 0 && (!enc_offsets || !encoded_lengths || !last_tokens)))
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_greedy_decode_resume: bad arguments");
    int64_t T = 0;
    for (int b = 0; b < B; ++b) {
        if (encoded_lengths[b] < 0 || encoded_lengths[b] > 0x7fffffff) return fail(c, AMIRA_ERR_INVALID_VALUE, "encoded_lengths out of range");
        T = encoded_lengths[b] > T ? encoded_lengths[b] : T;
    }
    return greedy_common(c, encoder_outputs, B, (int32_t)T, encoded_lengths, nullptr, states_1, states_2, tokens, n_tokens, n_steps,
                         enc_offsets, last_tokens);
    API_END(c)
}

// ---------------------------------------------------------------------------------------------- stream slots
int32_t amira_stream_open(amira_ctx *c, int32_t *slot) {
    API_BEGIN(c)
    if (!slot) return fail(c, AMIRA_ERR_INVALID_VALUE, "null slot pointer");
    for (size_t s = 0; s < c->slot_used.size(); ++s)
        if (!c->slot_used[s]) {
            const size_t off = s * 2 * kH;
            CK(cudaMemsetAsync(c->slot_s1 + off, 0, sizeof(float) * 2 * kH, c->stream), "slot reset");
            CK(cudaMemsetAsync(c->slot_s2 + off, 0, sizeof(float) * 2 * kH, c->stream), "slot reset");
            c->slot_used[s] = 1;
            *slot = (int32_t)s;
            return AMIRA_OK;
        }
    return fail(c, AMIRA_ERR_OUT_OF_MEMORY, "all stream slots in use (amira_config.max_streams)");
    API_END(c)
}

int32_t amira_stream_close(amira_ctx *c, int32_t slot) {
    API_BEGIN(c)
    if (slot < 0 || (size_t)slot >= c->slot_used.size() || !c->slot_used[(size_t)slot])
        return fail(c, AMIRA_ERR_INVALID_VALUE, "stream slot not open");
    c->slot_used[(size_t)slot] = 0;
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_stream_get_state(amira_ctx *c, int32_t slot, float *states_1, float *states_2) {
    API_BEGIN(c)
    if (slot < 0 || (size_t)slot >= c->slot_used.size() || !c->slot_used[(size_t)slot] || !states_1 || !states_2)
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_stream_get_state: bad arguments");
    const size_t off = (size_t)slot * 2 * kH;
    CK(cudaMemcpyAsync(states_1, c->slot_s1 + off, sizeof(float) * 2 * kH, cudaMemcpyDefault, c->stream), "state D2H");
    CK(cudaMemcpyAsync(states_2, c->slot_s2 + off, sizeof(float) * 2 * kH, cudaMemcpyDefault, c->stream), "state D2H");
    CK(cudaStreamSynchronize(c->stream), "state sync");
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_stream_set_state(amira_ctx *c, int32_t slot, const float *states_1, const float *states_2) {
    API_BEGIN(c)
    if (slot < 0 || (size_t)slot >= c->slot_used.size() || !c->slot_used[(size_t)slot] || !states_1 || !states_2)
        return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_stream_set_state: bad arguments");
    const size_t off = (size_t)slot * 2 * kH;
    CK(cudaMemcpyAsync(c->slot_s1 + off, states_1, sizeof(float) * 2 * kH, cudaMemcpyDefault, c->stream), "state H2D");
    CK(cudaMemcpyAsync(c->slot_s2 + off, states_2, sizeof(float) * 2 * kH, cudaMemcpyDefault, c->stream), "state H2D");
    CK(cudaStreamSynchronize(c->stream), "state sync");
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_stream_decode(amira_ctx *c, const int32_t *slots, int32_t n, const float *encoder_outputs, int32_t T,
                            const int64_t *encoded_lengths, int32_t *tokens, int32_t *n_tokens, int32_t *n_steps) {
    API_BEGIN(c)
    if (!slots && n > 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "null slots");
    return greedy_common(c, encoder_outputs, n, T, encoded_lengths, slots, nullptr, nullptr, tokens, n_tokens, n_steps);
    API_END(c)
}

// ---------------------------------------------------------------------------------------------- device-resident hand-off
// (the reference's CUDA shared-memory regions, src/cuda/cuda_helper.cu:63-183)
int32_t amira_device_alloc(amira_ctx *c, size_t bytes, void **dev_ptr) {
    API_BEGIN(c)
    if (!dev_ptr || bytes == 0) return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_device_alloc: bad arguments");
    *dev_ptr = nullptr;
    CK(cudaMalloc(dev_ptr, bytes), "device region alloc");
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_device_free(amira_ctx *c, void *dev_ptr) {
    API_BEGIN(c)
    if (!dev_ptr) return AMIRA_OK;
    CK(cudaStreamSynchronize(c->stream), "device region free sync");
    CK(cudaFree(dev_ptr), "device region free");
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_ipc_export(amira_ctx *c, const void *dev_ptr, amira_ipc_handle *handle) {
    API_BEGIN(c)
    static_assert(sizeof(amira_ipc_handle) == sizeof(cudaIpcMemHandle_t), "amira_ipc_handle is a cudaIpcMemHandle_t");
    if (!dev_ptr || !handle || !is_device_ptr(dev_ptr)) return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_ipc_export: not a device pointer");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)), "cudaIpcGetMemHandle");
    std::memcpy(handle->reserved, &h, sizeof(h));
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_ipc_import(amira_ctx *c, const amira_ipc_handle *handle, void **dev_ptr) {
    API_BEGIN(c)
    if (!handle || !dev_ptr) return fail(c, AMIRA_ERR_INVALID_VALUE, "amira_ipc_import: bad arguments");
    *dev_ptr = nullptr;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle->reserved, sizeof(h));
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return AMIRA_OK;
    API_END(c)
}

int32_t amira_ipc_close(amira_ctx *c, void *dev_ptr) {
    API_BEGIN(c)
    if (!dev_ptr) return AMIRA_OK;
    CK(cudaStreamSynchronize(c->stream), "ipc close sync");
    CK(cudaIpcCloseMemHandle(dev_ptr), "cudaIpcCloseMemHandle");
    return AMIRA_OK;
    API_END(c)
}

}  // extern "C"
