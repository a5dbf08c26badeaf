// common.h — internal declarations shared by the translation units of libamira_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "amira_b200.h"

namespace amira {

constexpr int kH = AMIRA_STATE_SIZE;      // 640
constexpr int kG = 4 * kH;                // 2560 gate rows
constexpr int kEnc = AMIRA_ENC_DIM;       // 1024
constexpr int kV = AMIRA_VOCAB_SIZE;      // 1030
constexpr int kEmbRows = 1025;
constexpr int kMel = AMIRA_N_MELS;        // 128
constexpr int kNfft = 512, kNbin = 257, kWin = 400, kHop = 160;

// ---- front-end tables (built on the host in double precision, tables.cpp) ----
// Mel filterbank for the kernel: filters in groups of four consecutive ones.  A group walks `steps` = its longest support;
// step s applies weight mel_w[woff + 4 s + j] (zero beyond filter j's own support) to power bin k0[j] + s.  Every table access is
// warp-uniform (the kernel maps a lane to a frame), the loop body is a dozen instructions, and the filters of a warp are a
// contiguous range of groups with about an eighth of the modelled cost each.
constexpr int kMelGroupsMax = 40, kMelWeightsMax = 1024, kFeWarps = 8;
struct alignas(16) MelGroup {
    int m0, nf, steps, woff;  // first filter, filters in the group (1..4), term steps, first weight
    int k0[4];                // first power bin of each filter
};
struct FrontendTables {
    float win[kNfft];               // Hann(400, symmetric) rounded to f32, centred in 512, zeros outside [56,456)
    int kstart[kMel];               // first non-zero FFT bin of each mel filter
    int kcnt[kMel];                 // number of non-zero bins
    int n_groups;
    int warp_group[kFeWarps + 1];   // groups of warp w: [warp_group[w], warp_group[w+1])
    alignas(16) MelGroup grp[kMelGroupsMax];
    alignas(16) float mel_w[kMelWeightsMax];  // woff is a multiple of 4: float4 loads
};
void build_frontend_tables(FrontendTables *t);
void build_mel_filterbank(float *fb /* [128][257] */);
void weights_random_init(float *blob, uint64_t seed, float blank_bias);

// ---- blob layout (element offsets) ----
struct BlobLayout {
    size_t emb, w_ih[2], w_hh[2], b_ih[2], b_hh[2], w_enc, b_enc, w_pred, b_pred, w_out, b_out, total;
};
BlobLayout blob_layout();

// a growable device / pinned-host scratch buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct TcWeights;  // decoder_tc.cu
// Device memory that a context and its forks (amira_ctx_fork: more submission lanes on the same GPU) share read-only: front-end
// tables, the fp32 weight blob, the derived weight tables and the split-bf16 tcgen05 operands.  Freed with the last context that
// holds it.  `version` counts weight loads: a lane refreshes its cached pointers when it falls behind.
struct SharedDev {
    int device = 0;
    FrontendTables *tables_dev = nullptr;
    float *w_blob = nullptr;
    float *g0p = nullptr, *whh0p = nullptr, *w1p = nullptr, *b1p = nullptr, *bjoint = nullptr, *woutp = nullptr, *boutp = nullptr;
    TcWeights *tc = nullptr;
    int coop_blocks_per_sm = 0;
    int version = 0;  // 0 = no weights loaded
    std::mutex mu;    // weight loads
    ~SharedDev();     // decoder.cu
};
struct DecoderPriv {  // per-context view of the shared weight tables + this context's decode workspace
    float *g0p = nullptr, *whh0p = nullptr, *w1p = nullptr, *b1p = nullptr, *bjoint = nullptr, *woutp = nullptr, *boutp = nullptr;
    DevBuf work;
    int coop_blocks_per_sm = 0;
    int *fail_count_dev = nullptr;  // valid after a greedy launch
    TcWeights *tc = nullptr;
    long long *ws_trace_dev = nullptr;  // decoder_ws.cu debug trace of the last launch (AMIRA_WS_TRACE=1)
};

// first-come-first-served turns of the batch-sized host -> device uploads of one GPU (decoder_tc.cu, api.cu)
std::mutex &upload_turn_mutex(int device);
int upload_fifo_mode();  // AMIRA_H2D_FIFO: 0 = off, 1 = the encoder-output uploads take turns (default; the front end's PCM uploads
                         // taking turns as well measured the same: 32.0 vs 32.1 ms per e2e step)

struct Ctx {
    amira_config cfg{};
    int device = 0;
    int sm_count = 0;
    // cudaFuncSetAttribute is per device: remembered per context, not per process (one process may hold contexts on several GPUs)
    bool attr_fe_i16 = false, attr_fe_f32 = false, attr_gemm = false;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream in use (own or caller-provided)
    std::mutex mu;
    std::string err;
    int64_t launches = 0;

    // optional per-kernel timing (amira_ctx_profile): CUDA events recorded on `stream` around named launches
    struct ProfSpan { int kernel; cudaEvent_t beg, end; };
    bool profiling = false;
    std::vector<ProfSpan> prof_spans;
    double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t prof_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    // front end; one set of scratch per in-flight chunk of a pipelined (host-buffer) call
    static constexpr int kMaxChunks = 8;
    FrontendTables *tables_dev = nullptr;
    DevBuf fe_meta[kMaxChunks];      // per chunk: starts[B], lens[B], tile prefix[B+1]
    DevBuf fe_partials[kMaxChunks];  // per tile x mel: (mean, M2) in double
    PinBuf fe_meta_pin[kMaxChunks];

    // host-buffer calls overlap H2D copies, kernels and D2H copies chunk by chunk: copies run on these two streams,
    // ordered against `stream` with events from this pool
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_pool[4 * kMaxChunks] = {};
    cudaEvent_t ev_block = nullptr;  // blocking-sync event: batch-sized calls sleep instead of spinning while the GPU works

    // staging for host-pointer arguments (slot per argument position)
    DevBuf stage[10];
    PinBuf pin[10];

    // decoder
    std::shared_ptr<SharedDev> shared;  // owner of tables_dev, w_blob and the weight tables behind `dec`
    int weights_version = 0;            // version of `shared` this context's cached pointers belong to
    bool has_weights = false;
    float *w_blob = nullptr;  // fp32 blob as loaded (owned by `shared`)
    DecoderPriv *dec = nullptr;

    // stream slots (WebSocket path): device-resident LSTM state
    float *slot_s1 = nullptr, *slot_s2 = nullptr;  // [max_streams][2][640]
    std::vector<uint8_t> slot_used;
};

// kernel ids for amira_ctx_kernel_ms
enum ProfKernel { PK_FE_LOGMEL = 0, PK_FE_NORMALIZE = 1, PK_ENC_PROJ = 2, PK_GREEDY = 3, PK_BYTES = 4, PK_COUNT = 5 };
struct ProfScope {  // RAII: records begin/end events around one launch when profiling is on
    Ctx *c; int idx;
    ProfScope(Ctx *ctx, int kernel) : c(ctx), idx(-1) {
        if (!c->profiling) return;
        Ctx::ProfSpan sp{kernel, nullptr, nullptr};
        if (cudaEventCreate(&sp.beg) != cudaSuccess || cudaEventCreate(&sp.end) != cudaSuccess) return;
        cudaEventRecord(sp.beg, c->stream);
        c->prof_spans.push_back(sp);
        idx = (int)c->prof_spans.size() - 1;
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof_spans[(size_t)idx].end, c->stream); }
};

// frontend.cu --------------------------------------------------------------------------------------------------
// wave_dev: int16 PCM or float samples; utterance b occupies [starts[b], starts[b]+lens[b]) (element units).
// features_dev [B][128][t_stride]; frames >= features_len are zero.  Packed layout: t_stride == 0 and foff_host[b] = first
// element of utterance b's [128][features_len_b] block relative to features_dev.
cudaError_t launch_frontend(Ctx *c, const void *wave_dev, bool is_pcm16, const int64_t *starts_host,
                            const int64_t *lens_host, int B, float *features_dev, int64_t t_stride, int slot = 0, int phase = 0,
                            const int64_t *foff_host = nullptr, bool normalize = true);
cudaError_t launch_bytes_to_f32(Ctx *c, const uint8_t *bytes_dev, size_t n_bytes, bool drop_odd, float *out_dev);
cudaError_t frontend_upload_tables(const FrontendTables *t);  // mel filterbank -> constant memory of the current device

// decoder.cu ---------------------------------------------------------------------------------------------------
cudaError_t decoder_prepare_weights(Ctx *c);  // derived tables from c->w_blob, published in c->shared
void decoder_adopt_shared(Ctx *c);            // refresh this context's cached weight pointers from c->shared (after a load / fork)
void decoder_release(Ctx *c);                 // this context's workspace only; the weights go with the last holder of c->shared
// enc_dev [B][1024][T]; lens_dev int32[B]; slots_dev nullable: when given, states live in c->slot_s1/2 rows
// slots[b] ([slot][2][640]); otherwise s1/s2 are [2][B][640] in/out (nullable => zero start, result dropped).
// enc_host != nullptr: the encoder outputs still live in host memory; the launcher uploads them into enc_dev chunk by
// chunk on c->h2d_stream and overlaps the upload with the encoder projection of the chunks already on the device.
cudaError_t launch_greedy_decode(Ctx *c, const float *enc_dev, const float *enc_host, int B, int T, const int32_t *lens_dev,
                                 const int32_t *lens_host, const int32_t *slots_dev, float *s1_dev, float *s2_dev, int32_t *tokens_dev,
                                 int32_t *ntok_dev, int32_t *nsteps_dev, const int64_t *enc_off_host = nullptr, int32_t *last_dev = nullptr);
const int32_t *decoder_fail_count_dev(Ctx *c);  // failed-stream counter of the last greedy launch
cudaError_t launch_decoder_joint(Ctx *c, const float *enc_dev, int B, int T, const int32_t *targets_dev, int U,
                                 const int32_t *tlen_dev, const float *in_s1, const float *in_s2, float *outputs,
                                 int32_t *prednet_lengths, float *out_s1, float *out_s2, int32_t *err_flag_dev);


// decoder_tc.cu (tcgen05 paths) --------------------------------------------------------------------------------
cudaError_t decoder_tc_prepare_weights(Ctx *c);
void decoder_tc_free(TcWeights *w);
cudaError_t launch_greedy_decode_tc(Ctx *c, const float *enc_dev, const float *enc_host, int B, int T, const int32_t *lens_dev,
                                    const int32_t *lens_host, const int32_t *slots_dev, float *s1_dev, float *s2_dev,
                                    int32_t *tokens_dev, int32_t *ntok_dev, int32_t *nsteps_dev, const int64_t *enc_off_host = nullptr,
                                    int32_t *last_dev = nullptr);
cudaError_t launch_tc_gemm(Ctx *c, const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_lo, const __nv_bfloat16 *w_hi,
                           const __nv_bfloat16 *w_lo, const float *bias, float *C, long long ldc, int M, int N, int K);
cudaError_t launch_split_rows(Ctx *c, const float *x, size_t ldx, __nv_bfloat16 *hi, __nv_bfloat16 *lo, size_t ldo,
                              size_t rows, size_t cols);
cudaError_t launch_split_transpose_enc(Ctx *c, const float *enc, int B, int T, const int *lens_dev, const int *eoff_dev,
                                       int row_base, __nv_bfloat16 *hi, __nv_bfloat16 *lo, const long long *src_off_dev = nullptr);


// decoder_ws.cu (weight-stationary dataflow engine, decode_engine = 4) ------------------------------------------
bool decoder_ws_supported(const Ctx *c);
cudaError_t decoder_ws_prepare(Ctx *c, TcWeights *w);
// The lane plan of a batch (decoder_ws.cu): MT M-tiles of 128 lanes share the batch's streams; rowinfo lists the streams with
// frames, longest first: the first Mpad start in the lanes, the others are taken off the queue when a lane's stream ends.
struct WsPlan {
    int MT = 1, Mpad = 128, n_streams = 0;
    int e_rows = 0;  // rows of E the streams cover (max over streams of first row + length)
    int spec = 0;  // ticks of an M-tile overlap by blank speculation from the start
};
// fills rowinfo [B] {stream, encoded length, first row of E, 0} (host memory)
WsPlan ws_plan_lanes(const int32_t *lens, const int *eoff, int B, int4 *rowinfo);
// work == nullptr: size query (*work_bytes receives the workspace size for plan.MT M-tiles).  E [sum of lengths][640] and the
// device copy of the plan's rowinfo from the caller.
cudaError_t launch_greedy_ws(Ctx *c, const float *E, int B, const WsPlan &plan, int T, const int4 *rowinfo_dev,
                             const int32_t *slots_dev, float *s1_dev, float *s2_dev, int32_t *tokens_dev, int32_t *ntok_dev,
                             int32_t *nsteps_dev, char *work, size_t *work_bytes, int32_t *last_dev = nullptr);

}  // namespace amira
