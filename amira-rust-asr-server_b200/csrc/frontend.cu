// frontend.cu — the `preprocessor` stage on B200 (sm_100a).
//
// Replaces, for B utterances per launch, what the reference does with one Triton round trip per utterance:
//   convert_audio (src/asr/pipeline.rs:127-139 -> src/performance_opts.rs:14-31)  i16 LE PCM -> f32
//   PreprocessorModel::infer_zero_copy (src/triton/model.rs:71-160)               waveform -> [1,128,T'] features
// Spec of the (absent, LFS) preprocessor ONNX: SURVEY.md 8(c).
//
// ONE kernel (fe_fused_kernel), persistent CTAs that pull 32-frame tiles from an atomic counter (longest utterances first):
//   global i16/f32 -> smem (128-bit loads, the next tile's samples prefetched with cp.async) -> pre-emphasis + reflect padding
//   staged in smem (exact: for PCM the pre-emphasised sample 100*s[n]-97*s[n-1] is an integer < 2^24) -> a half-warp per frame:
//   Hann window -> 512-point real FFT as a 256-point complex FFT factored 16 x 16 (radix-16 in registers, twiddle, 16x16
//   transpose through shared memory, radix-16 in registers) -> real-input split against the mirrored bin, whose values sit in
//   the partner lane (16 - lane) and come over with warp shuffles -> |X|^2 of the tile's 32 frames in shared memory;
//   then the mel reduction with a LANE PER FRAME (filters in groups of four, warp-uniform table loads, each power bin one
//   conflict-free shared-memory load) -> log -> tile of un-normalised log-mel + per-tile (mean, M2) partials -> coalesced
//   stores.  The CTA that completes the LAST tile of an utterance (atomic counter) merges the partials (Chan, fp64) and
//   normalises the utterance in place while its features are still in L2: (x - mean) / (std + 1e-5), frames >= features_len
//   zeroed — no second kernel, no second pass over HBM.
//   The FFT runs in fp64: an fp32 FFT leaves ~2e-4 max-abs error after normalisation on low mel bins (deep
//   fades under pre-emphasis), above the 1e-4 contract; B200 has a 1:2 fp64 pipe (DESIGN.md "front end").
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"

namespace amira {

namespace {

constexpr int TF = 32;                             // frames per tile
constexpr int FE_THREADS = 256;                    // 8 warps = 16 half-warps: half of a 32-frame tile is transformed at a time
constexpr int FE_WARPS = FE_THREADS / 32;
static_assert(FE_WARPS == kFeWarps, "the mel tables are split for this many warps");
constexpr int SPAN = (TF - 1) * kHop + kNfft;      // 5472 padded samples per tile
constexpr int RAW_CAP = SPAN + 16;                 // raw samples staged per tile (+ previous sample, alignment slack)
constexpr int kMelWeightsSmem = 704;                // floats of FrontendTables::mel_w kept in shared memory (>= the table in use)
constexpr int P_LD = kNbin;                        // power spectrum row stride (257: odd, so a lane per frame is conflict-free)
constexpr int OUT_LD = TF + 1;
constexpr int FE_HALVES = 2 * FE_WARPS;            // a half-warp (16 lanes) transforms one frame
constexpr int TLD = 17;                            // row stride (complex doubles) of the 16x16 transpose buffer
constexpr int TBUF = 16 * TLD;                     // doubles per half-warp exchange buffer (real and imaginary parts cross in turn)

struct FeMeta {
    const int64_t *starts;   // [B] first element of each utterance
    const int64_t *lens;     // [B] samples
    const int32_t *tile_pfx; // [B+1] first tile of each utterance in the launch's tile order (+ its tile count, see tile_cnt)
    const int32_t *tile_cnt; // [B] tiles of each utterance
    const int32_t *tile_b;   // [n_tiles] utterance of every tile (utterances in descending length order)
    const int64_t *foff;     // [B] first element of each utterance's [128][ld] feature block
    int32_t *done;           // [B] tiles finished per utterance (zero at launch); [B] = the tile counter
    int B;
    int n_tiles;
    int debug;               // bit 0: leave the log-mel un-normalised (amira_logmel_pcm16_packed); AMIRA_FE_DEBUG adds bits for
                             // timing attribution only: 2 skip the mel phase, 8 skip the transforms
};

__device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t n) {
    if (n <= 1) return 0;
    const int64_t p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - i;
}

// L2 residency hints.  The un-normalised log-mel of an utterance is written, and read back once for the normalisation a few
// tens of microseconds later: those stores ask L2 to keep the lines (evict_last), the streamed PCM and the final normalised
// features ask to go first (evict_first), so that the read-back is served from L2 instead of DRAM (ncu: 1.33 GB read for
// 0.57 GB of PCM before).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_hint_f32(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_f32x4(float4 *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

struct cplx {
    double x, y;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx mul_mi(cplx a) { return {a.y, -a.x}; }  // a * (-i)

// ---- 16-point DFT in registers (two radix-4 passes; forward transform, W = exp(-2 pi i / 16)), natural order in/out ----
__device__ __forceinline__ void radix4(cplx &x0, cplx &x1, cplx &x2, cplx &x3) {
    const cplx s0 = cadd(x0, x2), d0 = csub(x0, x2), s1 = cadd(x1, x3), d1 = mul_mi(csub(x1, x3));
    x0 = cadd(s0, s1);
    x2 = csub(s0, s1);
    x1 = cadd(d0, d1);
    x3 = csub(d0, d1);
}
__device__ __forceinline__ void dft16(cplx (&a)[16]) {
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, r = 0.70710678118654752440;
    // pass 1: for each n2, DFT over n1 of a[n2 + 4 n1]  ->  y[n2][k1] stored at a[n2 + 4 k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) radix4(a[n2], a[n2 + 4], a[n2 + 8], a[n2 + 12]);
    // twiddles W16^(n2 k1)
    a[1 + 4] = cmul(a[1 + 4], cplx{c1, -s1});    // W^1
    a[1 + 8] = cplx{r * (a[1 + 8].x + a[1 + 8].y), r * (a[1 + 8].y - a[1 + 8].x)};  // W^2
    a[1 + 12] = cmul(a[1 + 12], cplx{s1, -c1});  // W^3
    a[2 + 4] = cplx{r * (a[2 + 4].x + a[2 + 4].y), r * (a[2 + 4].y - a[2 + 4].x)};  // W^2
    a[2 + 8] = mul_mi(a[2 + 8]);                 // W^4
    a[2 + 12] = cplx{r * (a[2 + 12].y - a[2 + 12].x), -r * (a[2 + 12].x + a[2 + 12].y)};  // W^6
    a[3 + 4] = cmul(a[3 + 4], cplx{s1, -c1});    // W^3
    a[3 + 8] = cplx{r * (a[3 + 8].y - a[3 + 8].x), -r * (a[3 + 8].x + a[3 + 8].y)};      // W^6
    a[3 + 12] = cmul(a[3 + 12], cplx{-c1, s1});  // W^9
    // pass 2: for each k1, DFT over n2 of y[n2][k1]  ->  X[k1 + 4 k2] left at a[k2 + 4 k1]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) radix4(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
    // a[k2 + 4 k1] holds X[k1 + 4 k2]: transpose the 4x4 index grid back to natural order
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            const cplx t = a[i + 4 * j];
            a[i + 4 * j] = a[j + 4 * i];
            a[j + 4 * i] = t;
        }
}

template <typename RawT>
struct Stage;  // staged (pre-emphasised, reflect-padded) sample type per input type
template <>
struct Stage<int16_t> {
    using T = float;  // exact integers 100*s[n] - 97*s[n-1], |.| < 2^24
    static constexpr double kScale = 1.0 / (100.0 * 32768.0);
    __device__ static T make(int16_t cur, int16_t prev, bool first) {
        return (float)(100 * (int)cur - (first ? 0 : 97 * (int)prev));
    }
};
template <>
struct Stage<float> {
    using T = double;
    static constexpr double kScale = 1.0;
    __device__ static T make(float cur, float prev, bool first) {
        return first ? (double)cur : (double)cur - 0.97 * (double)prev;
    }
};

// (x - mu) * inv in place over one feature row of `ld` floats, frames >= L zeroed; one warp per row.  The row was written by other
// SMs moments ago: loads bypass L1 (ld.global.cg).  Scalar head up to the first 16-byte boundary (rows of the ragged layout start
// anywhere), 128-bit body with eight loads in flight per lane, scalar tail.
__device__ __forceinline__ void normalize_row(float *p, int64_t ld, int64_t L, float mu, float inv, int lane, uint64_t pol) {
    const int64_t head = min(ld, (int64_t)(((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) / sizeof(float)));
    if (lane < head) st_hint_f32(p + lane, lane < L ? (__ldcg(p + lane) - mu) * inv : 0.f, pol);
    float4 *p4 = reinterpret_cast<float4 *>(p + head);
    const int64_t n4 = (ld - head) / 4;
    constexpr int NU = 8;  // independent 128-bit loads in flight per lane (the CTA is alone on this utterance: latency-bound)
    for (int64_t i0 = lane; i0 < n4; i0 += 32 * NU) {
        float4 v[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int64_t i = i0 + 32 * u;
            v[u] = (i < n4 && head + i * 4 < L) ? __ldcg(p4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int64_t i = i0 + 32 * u, t = head + i * 4;
            if (i >= n4) continue;
            float4 w = v[u];
            if (t + 3 < L) {
                w.x = (w.x - mu) * inv; w.y = (w.y - mu) * inv; w.z = (w.z - mu) * inv; w.w = (w.w - mu) * inv;
            } else if (t >= L) {
                w = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                w.x = (w.x - mu) * inv;
                w.y = t + 1 < L ? (w.y - mu) * inv : 0.f;
                w.z = t + 2 < L ? (w.z - mu) * inv : 0.f;
                w.w = 0.f;
            }
            st_hint_f32x4(p4 + i, w, pol);
        }
    }
    const int64_t t_tail = head + n4 * 4 + lane;
    if (t_tail < ld) st_hint_f32(p + t_tail, t_tail < L ? (__ldcg(p + t_tail) - mu) * inv : 0.f, pol);
}

template <typename RawT>
__global__ void __launch_bounds__(FE_THREADS, 2)
fe_fused_kernel(const RawT *__restrict__ wave, FeMeta meta, const FrontendTables *__restrict__ tab,
                float *__restrict__ features, int64_t t_stride, double2 *__restrict__ partials) {
    using StT = typename Stage<RawT>::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xbuf = reinterpret_cast<double *>(smem_raw);                              // [FE_HALVES][TBUF] transpose exchange
    double2 *tw256 = reinterpret_cast<double2 *>(xbuf + FE_HALVES * TBUF);            // [16][16] W256^(m2 k1) at [k1][m2]
    double2 *tw512 = tw256 + 256;                                                     // [129] exp(-2 pi i k / 512)
    double2 *winp = tw512 + 130;                                                      // [256] scaled window pairs (w[2m], w[2m+1])
    StT *ystage = reinterpret_cast<StT *>(winp + 256);                                // [SPAN]; dead after the transforms ...
    float *outt = reinterpret_cast<float *>(ystage);                                  // ... [128][OUT_LD] log-mel tile aliases it
    float *pw = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(ystage) + sizeof(StT) * SPAN);  // [TF][P_LD] power spectra
    RawT *raw = reinterpret_cast<RawT *>(xbuf);  // [RAW_CAP] raw samples are staged before, the exchange buffers used after, the barrier
    // 16-bit input only: a second staging buffer that does NOT alias the exchange buffers, filled with cp.async for the NEXT
    // tile of this CTA while the current tile's frames are transformed
    constexpr bool kPrefetch = sizeof(RawT) == 2;
    RawT *raw2 = reinterpret_cast<RawT *>(pw + TF * P_LD);  // [RAW_CAP], 16-byte aligned
    constexpr size_t kRaw2Bytes = kPrefetch ? ((RAW_CAP * sizeof(RawT) + 15) & ~(size_t)15) : 0;
    // the grouped mel tables (FrontendTables::grp / mel_w / warp_group), copied once per CTA: walked by every warp for every tile
    int4 *s_grp = reinterpret_cast<int4 *>(reinterpret_cast<unsigned char *>(raw2) + kRaw2Bytes);   // [n_groups][2]
    float4 *s_melw = reinterpret_cast<float4 *>(s_grp + 2 * kMelGroupsMax);                         // [kMelWeightsSmem / 4]
    __shared__ int s_tile[2];   // tile being processed / next tile of this CTA
    __shared__ int s_last;      // this CTA finished the last tile of the utterance
    __shared__ int s_wg[kFeWarps + 1];
    float *s_mu = reinterpret_cast<float *>(ystage), *s_inv = s_mu + kMel;  // normalisation runs while the staging buffer is idle
    static_assert(sizeof(float) * kMel * OUT_LD <= sizeof(StT) * SPAN, "the log-mel tile must fit the staging buffer it aliases");
    static_assert((TF * P_LD * 4) % 16 == 0 && (SPAN * 4) % 16 == 0 && (FE_HALVES * TBUF * 8) % 16 == 0, "alignment of the carve-up");
    static_assert(sizeof(RawT) * RAW_CAP <= sizeof(double) * FE_HALVES * TBUF, "raw staging aliases the exchange buffers");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, hl = lane & 15;  // half-warp and lane within it
    const uint64_t pol_first = l2_policy_evict_first();
    // un-normalised tiles are read back by the normalisation: keep them (unless the un-normalised log-mel IS the output); the
    // normalisation's own stores (evict_first) hand the lines back
    const uint64_t pol_tile = (meta.debug & 1) ? pol_first : l2_policy_evict_last();

    // ---- loop-invariant tables (shared memory) ----
    // the window carries the input scale and the 1/2 of the real-input split (E = (Z + conj Z') / 2, ...)
    for (int i = tid; i < 256; i += FE_THREADS) {
        double sn, cs;
        sincospi(-2.0 * (double)(((i >> 4) * (i & 15)) & 255) / 256.0, &sn, &cs);  // entry [k1][m2] = W256^(m2 k1): lanes read consecutively
        tw256[i] = make_double2(cs, sn);
        winp[i] = make_double2((double)tab->win[2 * i] * (0.5 * Stage<RawT>::kScale), (double)tab->win[2 * i + 1] * (0.5 * Stage<RawT>::kScale));
    }
    for (int i = tid; i < 129; i += FE_THREADS) {
        double sn, cs;
        sincospi(-2.0 * (double)i / 512.0, &sn, &cs);
        tw512[i] = make_double2(cs, sn);
    }
    for (int i = tid; i < 2 * kMelGroupsMax; i += FE_THREADS) s_grp[i] = reinterpret_cast<const int4 *>(tab->grp)[i];
    for (int i = tid; i < kMelWeightsSmem / 4; i += FE_THREADS) s_melw[i] = reinterpret_cast<const float4 *>(tab->mel_w)[i];
    if (tid <= kFeWarps) s_wg[tid] = tab->warp_group[tid];
    double *myx = xbuf + (warp * 2 + half) * TBUF;
    int32_t *tile_counter = meta.done + meta.B;

    struct TileLoc {
        int b, f0, nf, span, delta, nvec;
        int64_t n, L, i0, g_lo, g_hi, foff;
        const RawT *x;
        const int4 *src;
        bool fits, fast;
    };
    constexpr int PER = 16 / sizeof(RawT);
    auto locate = [&](int tile) -> TileLoc {  // (utterance, first frame) of a tile and the signal span it needs
        TileLoc t;
        t.b = meta.tile_b[tile];
        t.n = meta.lens[t.b];
        t.x = wave + meta.starts[t.b];
        t.foff = meta.foff[t.b];
        t.L = t.n / kHop + 1;
        t.f0 = (tile - meta.tile_pfx[t.b]) * TF;
        t.nf = (int)min((int64_t)TF, t.L - t.f0);
        t.span = (t.nf - 1) * kHop + kNfft;
        t.i0 = (int64_t)t.f0 * kHop - kNfft / 2;  // signal index of padded sample 0 of the tile
        t.g_lo = t.i0 - 1;
        t.g_hi = t.i0 + t.span;
        if (t.g_lo < 0) { t.g_lo = 0; t.g_hi = max(t.g_hi, (int64_t)(kNfft / 2 + 2)); }
        if (t.g_hi > t.n) { t.g_hi = t.n; t.g_lo = min(t.g_lo, t.n - (kNfft / 2 + 2)); }
        if (t.g_lo < 0) t.g_lo = 0;
        t.fits = (t.g_hi - t.g_lo) <= RAW_CAP;  // false only for pathological tiny signals with long reflections
        // common case: the tile lies well inside its utterance — copy the 16-byte aligned image of x[i0-1 ..]
        // (raw[q + delta] = x[i0 - 1 + q])
        t.fast = t.fits && t.i0 >= 1 + PER && t.i0 + t.span + PER <= t.n;
        t.delta = 0;
        t.src = nullptr;
        t.nvec = 0;
        if (t.fast) {
            const uintptr_t ga = reinterpret_cast<uintptr_t>(t.x + (t.i0 - 1));
            t.delta = (int)((ga & 15) / sizeof(RawT));
            t.src = reinterpret_cast<const int4 *>(ga & ~(uintptr_t)15);
            t.nvec = (t.span + 1 + t.delta + PER - 1) / PER;  // <= RAW_CAP / PER
        }
        return t;
    };
    auto prefetch = [&](const TileLoc &t) {  // asynchronous staging of a fast tile into raw2 (one commit group per call)
        if (t.fast) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(raw2);
            for (int v = tid; v < t.nvec; v += FE_THREADS)
                asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + 16u * (uint32_t)v), "l"(t.src + v), "l"(pol_first) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // utterances without samples have no tile: their rows of the padded layout are all padding
    if (t_stride > 0)
        for (int b = blockIdx.x; b < meta.B; b += gridDim.x)
            if (meta.tile_cnt[b] == 0) {
                float *ub = features + meta.foff[b];
                for (int64_t i = tid; i < (int64_t)kMel * t_stride; i += FE_THREADS) ub[i] = 0.f;
            }

    // ---- deferred normalisation: the CTA that finished the LAST tile of an utterance (atomic counter) normalises it, from L2 ----
    int pend_b = -1, pend_old = 0;
    int64_t pend_L = 0, pend_ld = 0, pend_foff = 0;
    auto normalize_pending = [&]() {  // all threads; s_last was set by thread 0 and published by a barrier
        if (s_last && !(meta.debug & 1)) {
            __threadfence();  // acquire side of the counter: every other CTA's stores of this utterance are visible
            if (tid < kMel) {  // Chan et al. merge of the per-tile (count, mean, M2), one thread per mel row
                const int m = tid, t0 = meta.tile_pfx[pend_b], nt = meta.tile_cnt[pend_b];
                double cn = 0.0, cmean = 0.0, cm2 = 0.0;
                for (int t = 0; t < nt; ++t) {
                    const double2 q = __ldcg(partials + (size_t)(t0 + t) * kMel + m);
                    const double n2 = (double)min((int64_t)TF, pend_L - (int64_t)t * TF), nt2 = cn + n2, d = q.x - cmean;
                    cmean += d * (n2 / nt2);
                    cm2 += q.y + d * d * (cn * n2 / nt2);
                    cn = nt2;
                }
                const double sd = pend_L > 1 ? sqrt(cm2 / (double)(pend_L - 1)) : 0.0;
                s_mu[m] = (float)cmean;
                s_inv[m] = (float)(1.0 / (sd + 1e-5));
            }
            __syncthreads();
            float *ub = features + pend_foff;
            for (int m = warp; m < kMel; m += FE_WARPS) normalize_row(ub + (int64_t)m * pend_ld, pend_ld, pend_L, s_mu[m], s_inv[m], lane, pol_first);
            __syncthreads();  // s_mu / s_inv alias the staging buffer the caller writes next
        }
    };

    // ---- dynamic tile scheduler: every CTA holds the tile it works on and the one after it (whose samples are in flight) ----
    if (tid == 0) {
        s_tile[0] = atomicAdd(tile_counter, 1);
        s_tile[1] = atomicAdd(tile_counter, 1);
    }
    __syncthreads();  // also: tables visible
    int tile = s_tile[0], next_tile = s_tile[1];
    TileLoc cur;
    if (tile < meta.n_tiles) {
        cur = locate(tile);
        if (kPrefetch) prefetch(cur);
    }

    while (tile < meta.n_tiles) {
        const int b = cur.b, f0 = cur.f0, nf = cur.nf, span = cur.span;
        const int64_t n = cur.n, L = cur.L, i0 = cur.i0, g_lo = cur.g_lo, g_hi = cur.g_hi, foff_b = cur.foff;
        const RawT *x = cur.x;
        const bool fits = cur.fits, fast = cur.fast;
        const int delta = cur.delta;

        // the next tile's location: five dependent global loads, issued here so that their latency hides behind the staging below
        TileLoc nxt = cur;
        if (next_tile < meta.n_tiles) nxt = locate(next_tile);
        __syncthreads();  // previous tile fully consumed (ystage/outt/pw/raw reuse, s_tile read)
        if (tid == 0) {
            s_tile[0] = atomicAdd(tile_counter, 1);  // the tile after next; read after the next barrier
            s_last = pend_b >= 0 && pend_old == meta.tile_cnt[pend_b] - 1;  // the previous tile's completion count has arrived by now
        }
        // ---- stage the signal span [g_lo, g_hi) needed by this tile ----
        if (fast && kPrefetch) {
            asm volatile("cp.async.wait_all;" ::: "memory");  // this thread's share of the tile, requested during the previous tile
        } else if (fast) {
            int4 *dst = reinterpret_cast<int4 *>(raw);
#pragma unroll 4
            for (int v = tid; v < cur.nvec; v += FE_THREADS) dst[v] = __ldg(cur.src + v);
        } else if (fits) {
            const int cnt = (int)(g_hi - g_lo);
            // 128-bit path over the 16-byte aligned interior, scalar head/tail
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(x + g_lo);
            int head = (int)(((16 - (a0 & 15)) & 15) / sizeof(RawT));
            if (head > cnt) head = cnt;
            const int nvec = (cnt - head) / PER;
            for (int i = tid; i < head; i += FE_THREADS) raw[i] = x[g_lo + i];
            const int4 *src = reinterpret_cast<const int4 *>(x + g_lo + head);
            for (int v = tid; v < nvec; v += FE_THREADS) {
                const int4 q = __ldg(src + v);
                alignas(16) RawT tmp[PER];
                *reinterpret_cast<int4 *>(tmp) = q;
#pragma unroll
                for (int e = 0; e < PER; ++e) raw[head + v * PER + e] = tmp[e];
            }
            for (int i = head + nvec * PER + tid; i < cnt; i += FE_THREADS) raw[i] = x[g_lo + i];
        }
        __syncthreads();
        const int after_next = s_tile[0];
        normalize_pending();  // the utterance whose last tile this CTA finished in its previous iteration (rare: once per utterance)
        // ---- pre-emphasis (y[0] = x[0]; y[i] = x[i] - 0.97 x[i-1]) + reflect padding ----
        const bool interior = fits && i0 >= 1 && i0 + span <= n;
        if (interior) {
            // rq[q] = x[i0 - 1 + q]; two consecutive samples per thread and step (one staged pair store), unrolled for ILP
            const RawT *rq = ((fast && kPrefetch) ? raw2 : raw) + delta;
#pragma unroll 4
            for (int p = 2 * tid; p < span; p += 2 * FE_THREADS) {
                const RawT x0 = rq[p], x1 = rq[p + 1], x2 = rq[p + 2];
                const StT ya = Stage<RawT>::make(x1, x0, false), yb = Stage<RawT>::make(x2, x1, false);
                if constexpr (sizeof(StT) == 4) *reinterpret_cast<float2 *>(ystage + p) = make_float2((float)ya, (float)yb);
                else *reinterpret_cast<double2 *>(ystage + p) = make_double2((double)ya, (double)yb);
            }
        } else {
            for (int p = tid; p < span; p += FE_THREADS) {
                const int64_t r = reflect_index(i0 + p, n);
                RawT cur_s, prev = RawT(0);
                if (fits) {
                    cur_s = raw[r - g_lo];
                    if (r > 0) prev = raw[r - 1 - g_lo];
                } else {
                    cur_s = x[r];
                    if (r > 0) prev = x[r - 1];
                }
                ystage[p] = Stage<RawT>::make(cur_s, prev, r == 0);
            }
        }
        __syncthreads();
        // raw2 has been consumed: request the next tile of this CTA now, its loads fly during the transforms below
        if (next_tile < meta.n_tiles && kPrefetch) prefetch(nxt);

        // ---- a half-warp per frame, two frames per warp at a time: 512-point real FFT = 256-point complex FFT (z[m] = x[2m] +
        // i x[2m+1]) as 16 x 16: radix-16 in registers over m1 (m = 16 m1 + lane), twiddle W256^(lane k1), transpose through
        // shared memory, radix-16 over m2; then the real-input split against the mirrored bin and |X|^2 ----
        for (int fp = warp; 2 * fp < nf && !(meta.debug & 8); fp += FE_WARPS) {
            const int fl = 2 * fp + half;  // frames >= nf transform stale-but-finite staging data; their results are not used
            const StT *fr = ystage + min(fl, TF - 1) * kHop;
            cplx a[16];
            // the centred 400-sample window is zero outside [56, 456): the terms m1 = 0 and m1 = 15 vanish for every lane
            a[0] = cplx{0.0, 0.0};
            a[15] = cplx{0.0, 0.0};
#pragma unroll
            for (int m1 = 1; m1 < 15; ++m1) {
                const int m = 16 * m1 + hl;
                const double2 w = winp[m];
                a[m1].x = (double)fr[2 * m] * w.x;
                a[m1].y = (double)fr[2 * m + 1] * w.y;
            }
            dft16(a);  // Y[k1] = sum_m1 z[16 m1 + hl] W16^(m1 k1)
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) {
                const double2 t = tw256[k1 * 16 + hl];
                a[k1] = cmul(a[k1], cplx{t.x, t.y});
            }
            // 16 x 16 transpose through shared memory, real parts then imaginary parts (half the buffer of a complex exchange)
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) myx[k1 * TLD + hl] = a[k1].x;
            __syncwarp();
#pragma unroll
            for (int m2 = 0; m2 < 16; ++m2) a[m2].x = myx[hl * TLD + m2];
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) myx[k1 * TLD + hl] = a[k1].y;
            __syncwarp();
#pragma unroll
            for (int m2 = 0; m2 < 16; ++m2) a[m2].y = myx[hl * TLD + m2];
            dft16(a);  // Z[hl + 16 k2] = sum_m2 (...) W16^(m2 k2), already halved by the window scale
            __syncwarp();
            // X[k] = E + W512^k O, X[256-k] = conj(E - W512^k O) with E = Z[k] + conj Z[256-k], O = (Z[k] - conj Z[256-k]) / i.
            // Lane hl owns the pairs k = hl + 16 k2, k2 = 0..7; the mirrored bins 256 - k = (16 - hl) + 16 (15 - k2) are
            // registers 15 - k2 of the partner lane (16 - hl) & 15: eight complex values cross by warp shuffle.  Lane 0 is its
            // own partner with the registers shifted by one (256 - 16 k2 = 16 (16 - k2)); it also owns the self-paired bin 128.
            float *mypw = pw + min(fl, TF - 1) * P_LD;
            const int src_lane = (lane & 16) | ((16 - hl) & 15);
            cplx rv[8];
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                rv[k2].x = __shfl_sync(0xffffffffu, a[15 - k2].x, src_lane);
                rv[k2].y = __shfl_sync(0xffffffffu, a[15 - k2].y, src_lane);
            }
            const bool store = fl < nf;
#pragma unroll
            for (int k2 = 0; k2 <= 8; ++k2) {
                const int k = k2 < 8 ? hl + 16 * k2 : 128;
                cplx z = a[k2], zp;
                if (k2 == 8) zp = a[8];                                    // bin 128 (lane 0 only): its own mirror
                else if (k2 == 0) zp = hl == 0 ? a[0] : rv[0];             // bin 0 pairs with itself (-> X[0], X[256])
                else zp = hl == 0 ? rv[k2 - 1] : rv[k2];
                const double2 w = tw512[k];
                const double ex = z.x + zp.x, ey = z.y - zp.y;
                const double dx = z.x - zp.x, dy = z.y + zp.y;
                const double ox = dy, oy = -dx;                            // (Z[k] - conj Z[256-k]) / i
                const double wx = w.x * ox - w.y * oy, wy = w.x * oy + w.y * ox;
                const double px = ex + wx, py = ey + wy, qx = ex - wx, qy = ey - wy;
                if (store && (k2 < 8 || hl == 0)) {
                    mypw[k] = (float)(px * px + py * py);
                    if (k2 < 8) mypw[256 - k] = (float)(qx * qx + qy * qy);
                }
            }
            __syncwarp();
        }
        __syncthreads();  // all power spectra of the tile are in shared memory; the staging buffer is dead (outt may overwrite it)

        // ---- mel reduction, a lane per frame.  Filters come in groups of four with a common walk (FrontendTables::grp): every
        // table access is warp-uniform (one shared-memory wavefront, broadcast), every power bin one conflict-free shared-memory load (row
        // stride 257), four independent FFMA chains per step.  A rolled loop on purpose: the fully unrolled form (504 FFMA with
        // immediate weights) was instruction-fetch bound — each warp ran 2.4 KB of straight-line code once per tile ----
        if (!(meta.debug & 2)) {
            const float *prow = pw + lane * P_LD;  // lanes >= nf read stale-but-finite spectra; their results are not used
            const int g_end = s_wg[warp + 1];
            for (int g = s_wg[warp]; g < g_end; ++g) {
                const int4 ga = s_grp[2 * g];       // m0, nf, steps, woff
                const int4 gk = s_grp[2 * g + 1];   // k0[4]
                const float4 *wv = s_melw + (ga.w >> 2);
                const float *p0 = prow + gk.x, *p1 = prow + gk.y, *p2 = prow + gk.z, *p3 = prow + gk.w;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 2
                for (int st = 0; st < ga.z; ++st) {
                    const float4 w4 = wv[st];
                    a0 = fmaf(w4.x, p0[st], a0);
                    a1 = fmaf(w4.y, p1[st], a1);
                    a2 = fmaf(w4.z, p2[st], a2);
                    a3 = fmaf(w4.w, p3[st], a3);
                }
                float *o = outt + ga.x * OUT_LD + lane;
                o[0] = logf(a0 + 5.9604644775390625e-08f);  // log(mel + 2^-24)
                if (ga.y > 1) o[OUT_LD] = logf(a1 + 5.9604644775390625e-08f);
                if (ga.y > 2) o[2 * OUT_LD] = logf(a2 + 5.9604644775390625e-08f);
                if (ga.y > 3) o[3 * OUT_LD] = logf(a3 + 5.9604644775390625e-08f);
            }
        }
        __syncthreads();

        // ---- per-tile statistics (two threads per mel row, 16 frames each) + coalesced store of the tile ----
        {
            static_assert(FE_THREADS == 2 * kMel && TF == 32, "two threads per mel row");
            const int m = tid >> 1, fb0 = (tid & 1) * 16, fe = min(nf, fb0 + 16);
            // mean in fp64 (it is subtracted from values of magnitude ~10 whose spread may be 1e-2), M2 in fp32 relative to it
            const float *orow = outt + m * OUT_LD;
            double s0 = 0.0, s1 = 0.0;
            int f = fb0;
            for (; f + 1 < fe; f += 2) { s0 += (double)orow[f]; s1 += (double)orow[f + 1]; }
            if (f < fe) s0 += (double)orow[f];
            double sum = s0 + s1;
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            const double mean = sum / nf;
            const float mf = (float)mean, ml = (float)(mean - (double)mf);
            float q0 = 0.f, q1 = 0.f;
            for (f = fb0; f + 1 < fe; f += 2) {
                const float d0 = (orow[f] - mf) - ml, d1 = (orow[f + 1] - mf) - ml;
                q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
            }
            if (f < fe) { const float d = (orow[f] - mf) - ml; q0 = fmaf(d, d, q0); }
            float q = q0 + q1;
            q += __shfl_xor_sync(0xffffffffu, q, 1);
            if (!(tid & 1)) partials[(size_t)tile * kMel + m] = make_double2(mean, (double)q);
        }
        // row stride: t_stride (padded layout) or, packed (t_stride == 0), the utterance's own frame count
        const int64_t ld = t_stride > 0 ? t_stride : L;
        float *dst = features + foff_b + f0;
        for (int m = warp; m < kMel; m += FE_WARPS)
            if (lane < nf) st_hint_f32(dst + (size_t)m * ld + lane, outt[m * OUT_LD + lane], pol_tile);

        // ---- completion of the utterance: count this tile; the answer (am I the last?) is consumed during the NEXT tile, so the
        // atomic's round trip is off the per-tile critical path ----
        __syncthreads();  // every thread's feature and partial stores are ordered before thread 0's fence below (cumulativity)
        if (tid == 0) {
            __threadfence();
            pend_old = atomicAdd(meta.done + b, 1);
        }
        pend_b = b; pend_L = L; pend_ld = ld; pend_foff = foff_b;
        cur = nxt;
        tile = next_tile;
        next_tile = after_next;
    }
    // the last tile this CTA worked on
    __syncthreads();
    if (tid == 0) s_last = pend_b >= 0 && pend_old == meta.tile_cnt[pend_b] - 1;
    __syncthreads();
    normalize_pending();
}

__global__ void bytes_to_f32_kernel(const uint8_t *__restrict__ in, size_t n_bytes, int drop_odd, float *__restrict__ out) {
    // src/performance_opts.rs:14-31: LE i16 / 32768; an odd trailing byte b -> (b as i16) / 128 (zero-extended)
    const size_t n_pairs = n_bytes / 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (aligned) {
        const size_t nvec = n_pairs / 8;
        const int4 *src = reinterpret_cast<const int4 *>(in);
        float4 *dst = reinterpret_cast<float4 *>(out);
        for (size_t v = i0; v < nvec; v += stride) {
            const int4 q = __ldg(src + v);
            const int w[4] = {q.x, q.y, q.z, q.w};
            float f[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                f[2 * e] = (float)(int16_t)(w[e] & 0xffff) * (1.0f / 32768.0f);
                f[2 * e + 1] = (float)(int16_t)((unsigned)w[e] >> 16) * (1.0f / 32768.0f);
            }
            dst[2 * v] = make_float4(f[0], f[1], f[2], f[3]);
            dst[2 * v + 1] = make_float4(f[4], f[5], f[6], f[7]);
        }
        for (size_t i = nvec * 8 + i0; i < n_pairs; i += stride) {
            const int16_t s = (int16_t)((uint16_t)in[2 * i] | ((uint16_t)in[2 * i + 1] << 8));
            out[i] = (float)s * (1.0f / 32768.0f);
        }
    } else {
        for (size_t i = i0; i < n_pairs; i += stride) {
            const int16_t s = (int16_t)((uint16_t)in[2 * i] | ((uint16_t)in[2 * i + 1] << 8));
            out[i] = (float)s * (1.0f / 32768.0f);
        }
    }
    if (i0 == 0 && (n_bytes & 1) && !drop_odd) out[n_pairs] = (float)(int16_t)in[n_bytes - 1] / 128.0f;
}

template <typename RawT>
size_t fe_smem_bytes() {
    return sizeof(double) * FE_HALVES * TBUF + sizeof(double2) * (256 + 130 + 256) + sizeof(typename Stage<RawT>::T) * SPAN + sizeof(float) * TF * P_LD +
           (sizeof(RawT) == 2 ? ((RAW_CAP * sizeof(RawT) + 15) & ~(size_t)15) : 0) +  // raw2 (16-bit input: prefetched staging)
           sizeof(int4) * 2 * kMelGroupsMax + sizeof(float) * kMelWeightsSmem;         // grouped mel tables
}

}  // namespace

cudaError_t frontend_upload_tables(const FrontendTables *t) {  // sanity of the grouped mel tables (tables.cpp)
    int n = 0;
    for (int g = 0; g < t->n_groups; ++g) n += t->grp[g].nf;
    const MelGroup &last = t->grp[t->n_groups > 0 ? t->n_groups - 1 : 0];
    return (t->n_groups > 0 && t->n_groups <= kMelGroupsMax && n == kMel && t->warp_group[kFeWarps] == t->n_groups &&
            last.woff + 4 * last.steps <= kMelWeightsSmem) ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_frontend(Ctx *c, const void *wave_dev, bool is_pcm16, const int64_t *starts_host,
                            const int64_t *lens_host, int B, float *features_dev, int64_t t_stride, int slot, int phase,
                            const int64_t *foff_host, bool normalize) {
    if (B <= 0) return cudaSuccess;
    if (slot < 0 || slot >= Ctx::kMaxChunks) return cudaErrorInvalidValue;
    DevBuf &fe_meta = c->fe_meta[slot], &fe_partials = c->fe_partials[slot];
    PinBuf &fe_meta_pin = c->fe_meta_pin[slot];
    // host metadata: starts, lens, feature offsets, tile tables, zeroed completion counters -> one pinned block, one async copy
    int64_t tiles = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t L = lens_host[b] <= 0 ? 0 : lens_host[b] / kHop + 1;
        tiles += (L + TF - 1) / TF;
    }
    if (tiles > 0x7fffffff) return cudaErrorInvalidValue;
    const size_t meta_bytes = sizeof(int64_t) * 3 * (size_t)B + sizeof(int32_t) * (3 * (size_t)B + 2 + (size_t)tiles);
    cudaError_t e;
    if ((e = fe_meta_pin.reserve(meta_bytes)) != cudaSuccess) return e;
    if ((e = fe_meta.reserve(meta_bytes)) != cudaSuccess) return e;
    int64_t *h_starts = fe_meta_pin.as<int64_t>();
    int64_t *h_lens = h_starts + B;
    int64_t *h_foff = h_lens + B;
    int32_t *h_pfx = reinterpret_cast<int32_t *>(h_foff + B);
    int32_t *h_cnt = h_pfx + B;
    int32_t *h_done = h_cnt + B;         // [B] tiles finished + [1] tile counter, all zero at launch
    int32_t *h_tile_b = h_done + B + 1;  // one load instead of a binary search per tile
    // tiles are handed out in this order: longest utterances first, so the normalisation that trails the last tile of the
    // launch belongs to a short utterance
    std::vector<int32_t> order((size_t)B);
    for (int b = 0; b < B; ++b) order[(size_t)b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return lens_host[x] > lens_host[y]; });
    tiles = 0;
    for (int i = 0; i < B; ++i) {
        const int b = order[(size_t)i];
        h_starts[b] = starts_host[b];
        h_lens[b] = lens_host[b];
        h_foff[b] = foff_host ? foff_host[b] : (int64_t)b * kMel * t_stride;
        const int64_t L = lens_host[b] <= 0 ? 0 : lens_host[b] / kHop + 1;
        const int64_t nt = (L + TF - 1) / TF;
        h_pfx[b] = (int32_t)tiles;
        h_cnt[b] = (int32_t)nt;
        h_done[b] = 0;
        for (int64_t t = 0; t < nt; ++t) h_tile_b[tiles + t] = b;
        tiles += nt;
    }
    h_done[B] = 0;
    // phase 1 = metadata upload only, 2 = kernel only (metadata already uploaded for this slot), 0 = both.  Pipelined host
    // calls upload every chunk's metadata BEFORE queueing the bulk copies: a small copy that becomes runnable later would
    // wait in the copy engine behind all bulk copies already queued there, and its kernel with it.
    if (phase != 2 && (e = cudaMemcpyAsync(fe_meta.p, fe_meta_pin.p, meta_bytes, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess)
        return e;
    if ((e = fe_partials.reserve(sizeof(double2) * (size_t)std::max<int64_t>(tiles, 1) * kMel)) != cudaSuccess) return e;
    if (phase == 1) return cudaSuccess;
    FeMeta meta;
    meta.starts = fe_meta.as<int64_t>();
    meta.lens = meta.starts + B;
    meta.foff = meta.lens + B;
    meta.tile_pfx = reinterpret_cast<const int32_t *>(meta.foff + B);
    meta.tile_cnt = meta.tile_pfx + B;
    meta.done = reinterpret_cast<int32_t *>(fe_meta.as<int64_t>() + 3 * (size_t)B) + 2 * (size_t)B;
    meta.tile_b = meta.done + B + 1;
    meta.B = B;
    meta.n_tiles = (int)tiles;
    meta.debug = getenv("AMIRA_FE_DEBUG") ? atoi(getenv("AMIRA_FE_DEBUG")) : 0;
    if (!normalize) meta.debug |= 1;  // un-normalised log-mel (the incremental streaming path normalises with running statistics)

    ProfScope prof(c, PK_FE_LOGMEL);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)c->sm_count * 2));
    if (is_pcm16) {
        const size_t smem = fe_smem_bytes<int16_t>();
        if (!c->attr_fe_i16) {
            if ((e = cudaFuncSetAttribute(fe_fused_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            c->attr_fe_i16 = true;
        }
        fe_fused_kernel<int16_t><<<grid, FE_THREADS, smem, c->stream>>>(
            static_cast<const int16_t *>(wave_dev), meta, c->tables_dev, features_dev, t_stride, fe_partials.as<double2>());
    } else {
        const size_t smem = fe_smem_bytes<float>();
        if (!c->attr_fe_f32) {
            if ((e = cudaFuncSetAttribute(fe_fused_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            c->attr_fe_f32 = true;
        }
        fe_fused_kernel<float><<<grid, FE_THREADS, smem, c->stream>>>(
            static_cast<const float *>(wave_dev), meta, c->tables_dev, features_dev, t_stride, fe_partials.as<double2>());
    }
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_bytes_to_f32(Ctx *c, const uint8_t *bytes_dev, size_t n_bytes, bool drop_odd, float *out_dev) {
    if (n_bytes == 0) return cudaSuccess;
    const size_t work = n_bytes / 16 + 1;
    const int grid = (int)std::min<size_t>((work + 255) / 256, (size_t)c->sm_count * 8);
    ProfScope prof(c, PK_BYTES);
    bytes_to_f32_kernel<<<grid, 256, 0, c->stream>>>(bytes_dev, n_bytes, drop_odd ? 1 : 0, out_dev);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace amira
