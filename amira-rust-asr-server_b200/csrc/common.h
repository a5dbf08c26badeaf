// common.h — internal declarations shared by the translation units of libamira_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
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

// ---- front-end tables (built on the host in double precision, weights.cpp) ----
struct FrontendTables {
    float win[kNfft];     // Hann(400, symmetric) centred in 512, zeros outside [56,456)
    float2 tw[256];       // exp(-2*pi*i*e/512), e = 0..255
    int kstart[kMel];     // first non-zero FFT bin of each mel filter
    int kcnt[kMel];       // number of non-zero bins
    int woff[kMel];       // offset of the filter's weights in melw
    float melw[512];      // packed non-zero Slaney mel weights (504 used)
};
void build_frontend_tables(FrontendTables *t);
void weights_random_init(float *blob, uint64_t seed, float blank_bias);

// ---- blob layout (element offsets) ----
struct BlobLayout {
    size_t emb, w_ih[2], w_hh[2], b_ih[2], b_hh[2], w_enc, b_enc, w_pred, b_pred, w_out, b_out, total;
};
BlobLayout blob_layout();

struct Ctx;

// frontend.cu
cudaError_t launch_frontend(Ctx *c, const void *wave_dev, bool is_pcm16, const int64_t *starts_host,
                            const int64_t *lens_host, int B, int64_t total_elems, float *features_dev,
                            int64_t t_stride, int64_t *features_lens_host);
cudaError_t launch_bytes_to_f32(Ctx *c, const uint8_t *bytes_dev, size_t n_bytes, bool drop_odd, float *out_dev);

// decoder.cu
cudaError_t decoder_prepare_weights(Ctx *c);
cudaError_t launch_greedy_decode(Ctx *c, const float *enc_dev, int B, int T, const int64_t *lens_host,
                                 const int32_t *slots_dev /*nullable*/, float *s1_dev, float *s2_dev,
                                 int32_t *tokens_dev, int32_t *ntok_dev, int32_t *nsteps_dev);
cudaError_t launch_decoder_joint(Ctx *c, const float *enc_dev, int B, int T, const int32_t *targets_dev, int U,
                                 const int32_t *tlen_dev, const float *in_s1, const float *in_s2, float *outputs,
                                 int32_t *prednet_lengths, float *out_s1, float *out_s2, int32_t *err_flag_dev);

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
};

struct Ctx {
    amira_config cfg{};
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream in use (own or caller-provided)
    std::mutex mu;
    std::string err;
    int64_t launches = 0;

    // front end
    FrontendTables *tables_dev = nullptr;
    DevBuf fe_meta;      // starts / lens / tile prefix
    DevBuf fe_partials;  // per-tile (sum, M2)
    DevBuf in_stage, out_stage, aux_stage[6];
    PinBuf pin_in, pin_out;

    // decoder
    bool has_weights = false;
    float *w_blob = nullptr;  // fp32 blob as loaded
    DevBuf dec_derived;       // G1 table etc. (decoder.cu owns the layout)
    DevBuf dec_work;          // per-call workspace
    DevBuf dec_ctrl;
    void *dec_priv = nullptr; // decoder.cu private struct

    // stream slots
    float *slot_s1 = nullptr, *slot_s2 = nullptr;  // [max_streams][2][640]
    std::vector<uint8_t> slot_used;
};

}  // namespace amira
