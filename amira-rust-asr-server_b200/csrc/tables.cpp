// tables.cpp — host-side constant tables of libamira_b200.so: front-end window / mel filterbank, the weight
// blob layout and the seeded random-init stand-in for the absent decoder_joint ONNX weights
// (model-repo/decoder_joint/1/model.onnx is a Git-LFS pointer in the reference).
#include <cmath>
#include <cstring>

#include "amira_hann400.h"
#include "common.h"

namespace amira {

// Slaney mel scale (librosa.filters.mel(sr=16000, n_fft=512, n_mels=128, fmin=0, fmax=8000, norm="slaney"),
// the filterbank NeMo's AudioToMelSpectrogramPreprocessor stores).
static double hz2mel(double f) {
    const double f_sp = 200.0 / 3.0, brk = 1000.0, brk_mel = brk / f_sp, step = std::log(6.4) / 27.0;
    return f >= brk ? brk_mel + std::log(f / brk) / step : f / f_sp;
}
static double mel2hz(double m) {
    const double f_sp = 200.0 / 3.0, brk = 1000.0, brk_mel = brk / f_sp, step = std::log(6.4) / 27.0;
    return m >= brk_mel ? brk * std::exp(step * (m - brk_mel)) : f_sp * m;
}

void build_mel_filterbank(float *fb) {
    double edge[kMel + 2];
    const double lo = hz2mel(0.0), hi = hz2mel(8000.0);
    for (int i = 0; i < kMel + 2; ++i) edge[i] = mel2hz(lo + (hi - lo) * i / (kMel + 1));
    for (int m = 0; m < kMel; ++m) {
        const double rise = edge[m + 1] - edge[m], fall = edge[m + 2] - edge[m + 1];
        const double area_norm = 2.0 / (edge[m + 2] - edge[m]);
        for (int k = 0; k < kNbin; ++k) {
            const double f = 8000.0 * k / (kNbin - 1);
            double w = std::fmin((f - edge[m]) / rise, (edge[m + 2] - f) / fall);
            if (w < 0.0) w = 0.0;
            fb[m * kNbin + k] = (float)(w * area_norm);
        }
    }
}

void build_frontend_tables(FrontendTables *t) {
    std::memset(t, 0, sizeof(*t));
    // window = the float32 buffer torch.hann_window(400, periodic=False) yields (include/amira_hann400.h)
    for (int n = 0; n < kWin; ++n) std::memcpy(&t->win[(kNfft - kWin) / 2 + n], &AMIRA_HANN400_BITS[n], sizeof(float));
    std::vector<float> fb((size_t)kMel * kNbin);
    build_mel_filterbank(fb.data());
    for (int m = 0; m < kMel; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < kNbin; ++k)
            if (fb[m * kNbin + k] != 0.0f) {
                if (first < 0) first = k;
                last = k;
            }
        t->kstart[m] = first < 0 ? 0 : first;
        t->kcnt[m] = first < 0 ? 0 : last - first + 1;
    }
    // contiguous filter ranges per warp with equal modelled cost (2 per non-zero weight + 26 per filter: log + store), then
    // groups of four inside each range
    long total = 0;
    for (int m = 0; m < kMel; ++m) total += 2 * t->kcnt[m] + 26;
    int split[kFeWarps + 1];
    split[0] = 0;
    long run = 0;
    int m = 0;
    for (int w = 1; w < kFeWarps; ++w) {
        while (m < kMel && run + (2 * t->kcnt[m] + 26) / 2 < total * w / kFeWarps) run += 2 * t->kcnt[m++] + 26;
        split[w] = m;
    }
    split[kFeWarps] = kMel;
    int g = 0, woff = 0;
    for (int w = 0; w < kFeWarps; ++w) {
        t->warp_group[w] = g;
        for (int m0 = split[w]; m0 < split[w + 1]; m0 += 4) {
            if (g >= kMelGroupsMax) { std::fprintf(stderr, "amira_b200: mel group table overflow\n"); return; }
            MelGroup &G = t->grp[g];
            G.m0 = m0;
            G.nf = split[w + 1] - m0 < 4 ? split[w + 1] - m0 : 4;
            G.steps = 0;
            for (int j = 0; j < G.nf; ++j) G.steps = t->kcnt[m0 + j] > G.steps ? t->kcnt[m0 + j] : G.steps;
            G.woff = woff;
            for (int j = 0; j < 4; ++j) {
                const int mj = j < G.nf ? m0 + j : m0;
                // a filter shorter than the group's walk reads a few bins past its support with zero weights: keep them inside the row
                G.k0[j] = t->kstart[mj] + G.steps <= kNbin ? t->kstart[mj] : kNbin - G.steps;
            }
            if (woff + 4 * G.steps > kMelWeightsMax) { std::fprintf(stderr, "amira_b200: mel weight table overflow\n"); return; }
            for (int s2 = 0; s2 < G.steps; ++s2)
                for (int j = 0; j < 4; ++j) {
                    float wv = 0.0f;
                    if (j < G.nf) {
                        const int k = G.k0[j] + s2, r = k - t->kstart[m0 + j];
                        if (r >= 0 && r < t->kcnt[m0 + j]) wv = fb[(m0 + j) * kNbin + k];
                    }
                    t->mel_w[woff + 4 * s2 + j] = wv;
                }
            woff += 4 * G.steps;
            ++g;
        }
    }
    t->warp_group[kFeWarps] = g;
    t->n_groups = g;
}

BlobLayout blob_layout() {
    BlobLayout L{};
    size_t o = 0;
    L.emb = o;
    o += (size_t)kEmbRows * kH;
    for (int l = 0; l < 2; ++l) {
        L.w_ih[l] = o;
        o += (size_t)kG * kH;
        L.w_hh[l] = o;
        o += (size_t)kG * kH;
        L.b_ih[l] = o;
        o += kG;
        L.b_hh[l] = o;
        o += kG;
    }
    L.w_enc = o;
    o += (size_t)kH * kEnc;
    L.b_enc = o;
    o += kH;
    L.w_pred = o;
    o += (size_t)kH * kH;
    L.b_pred = o;
    o += kH;
    L.w_out = o;
    o += (size_t)kV * kH;
    L.b_out = o;
    o += kV;
    L.total = o;
    return L;
}

// Seeded stand-in weights.  Generator spec (shared with the test oracle so both sides can be seeded alike):
// SplitMix64 stream per tensor, seed + 1000003 * tensor_index (1-based, blob order); u01 = top 24 bits / 2^24;
// LSTM/Linear ~ U(-1/sqrt(fan), 1/sqrt(fan)) with fan = 640 (1024 for the encoder projection); embedding ~
// Irwin-Hall(12) - 6 with the blank row (1024) zero; blank_bias is added to b_out[1024].
namespace {
struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    float u01() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
};
void uniform_fill(float *p, size_t n, uint64_t seed, float bound) {
    SplitMix g{seed};
    for (size_t i = 0; i < n; ++i) p[i] = (2.0f * g.u01() - 1.0f) * bound;
}
void normal_fill(float *p, size_t n, uint64_t seed) {
    SplitMix g{seed};
    for (size_t i = 0; i < n; ++i) {
        float a = 0.0f;
        for (int j = 0; j < 12; ++j) a += g.u01();
        p[i] = a - 6.0f;
    }
}
}  // namespace

void weights_random_init(float *blob, uint64_t seed, float blank_bias) {
    const BlobLayout L = blob_layout();
    const float kh = 1.0f / std::sqrt((float)kH), ke = 1.0f / std::sqrt((float)kEnc);
    uint64_t idx = 0;
    auto next_seed = [&]() { return seed + 1000003ULL * (++idx); };
    normal_fill(blob + L.emb, (size_t)kEmbRows * kH, next_seed());
    std::memset(blob + L.emb + (size_t)AMIRA_BLANK_ID * kH, 0, sizeof(float) * kH);
    for (int l = 0; l < 2; ++l) {
        uniform_fill(blob + L.w_ih[l], (size_t)kG * kH, next_seed(), kh);
        uniform_fill(blob + L.w_hh[l], (size_t)kG * kH, next_seed(), kh);
        uniform_fill(blob + L.b_ih[l], kG, next_seed(), kh);
        uniform_fill(blob + L.b_hh[l], kG, next_seed(), kh);
    }
    uniform_fill(blob + L.w_enc, (size_t)kH * kEnc, next_seed(), ke);
    uniform_fill(blob + L.b_enc, kH, next_seed(), ke);
    uniform_fill(blob + L.w_pred, (size_t)kH * kH, next_seed(), kh);
    uniform_fill(blob + L.b_pred, kH, next_seed(), kh);
    uniform_fill(blob + L.w_out, (size_t)kV * kH, next_seed(), kh);
    uniform_fill(blob + L.b_out, kV, next_seed(), kh);
    blob[L.b_out + AMIRA_BLANK_ID] += blank_bias;
}

}  // namespace amira
