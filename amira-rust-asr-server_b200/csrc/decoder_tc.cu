// decoder_tc.cu — tcgen05 (5th-gen tensor core) paths of the decoder_joint stage on sm_100a.
//
//   tc_gemm_kernel        C[M][N] = A[M][K] W[N][K]^T (+bias) with split-bf16 operands (hi*hi + lo*hi + hi*lo, fp32
//                         accumulate in TMEM): TMA -> 128B-swizzled smem ring -> tcgen05.mma issued by one thread ->
//                         tcgen05.ld epilogue.  Used for the hoisted encoder projection E = enc W_enc^T + b
//                         (M = sum of encoder frames, the one large GEMM of the path) and as the unit-testable proof of
//                         the descriptor / swizzle conventions (tests/test_gpu_tcgen05.py).
//   split / transpose     fp32 -> (hi, lo) bf16 operand kernels; host side of the weight preparation and of the launch of the
//                         persistent decode kernel (decoder_ws.cu).
#include <cooperative_groups.h>
#include <cuda.h>
#include <mutex>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>

#include "common.h"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace amira {

// ------------------------------------------------------------------------------------------------ tensor maps
cudaError_t make_tmap_bf16(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                           uint32_t box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !p) return cudaErrorNotSupported;
        fn = reinterpret_cast<EncodeFn>(p);
    }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {row_stride_elems * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)tc::BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

namespace {

using namespace tc;

// ------------------------------------------------------------------------------------------------ split kernels
// x[rows][cols] fp32 (row stride ldx) -> hi/lo bf16 [rows][cols] (row stride ldo)
__global__ void split_rows_kernel(const float *__restrict__ x, size_t ldx, __nv_bfloat16 *__restrict__ hi,
                                  __nv_bfloat16 *__restrict__ lo, size_t ldo, size_t rows, size_t cols) {
    const size_t n = rows * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / cols, c = i % cols;
        __nv_bfloat16 h, l;
        split_bf16(x[r * ldx + c], h, l);
        hi[r * ldo + c] = h;
        lo[r * ldo + c] = l;
    }
}

// encoder_outputs [B][1024][T] (t contiguous) -> K-major rows [(eoff[b] + t)][1024] hi/lo bf16 (valid frames packed back to
// back: frames >= lens[b] are never projected), via a 128x32 smem transpose.  Packed input (src_off != nullptr): stream b is
// a [1024][lens[b]] block at enc + src_off[b].
__global__ void __launch_bounds__(256)
split_transpose_enc_kernel(const float *__restrict__ enc, int T, const int *__restrict__ lens, const int *__restrict__ eoff,
                           const long long *__restrict__ src_off, int row_base, __nv_bfloat16 *__restrict__ hi,
                           __nv_bfloat16 *__restrict__ lo) {
    // tile = 128 features x 32 frames: 128-byte reads along t, 128-byte writes (64 bf16 pairs) along f
    __shared__ float tile[128][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, f0 = blockIdx.y * 128;
    const int len = lens[b];
    if (t0 >= len) return;
    const size_t r0 = (size_t)(eoff[b] - row_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *src = src_off ? enc + src_off[b] : enc + (size_t)b * kEnc * T;
    const int ld = src_off ? len : T;
    const bool in = t0 + lane < ld;
#pragma unroll
    for (int r = warp; r < 128; r += 8) tile[r][lane] = in ? __ldg(src + (size_t)(f0 + r) * ld + t0 + lane) : 0.f;
    __syncthreads();
#pragma unroll
    for (int tt = warp; tt < 32; tt += 8) {
        const int t = t0 + tt;
        if (t >= len) continue;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int f = 64 * g + 2 * lane;
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(tile[f][tt], h0, l0);
            split_bf16(tile[f + 1][tt], h1, l1);
            const size_t o = (r0 + t) * kEnc + f0 + f;
            *reinterpret_cast<__nv_bfloat162 *>(hi + o) = __halves2bfloat162(h0, h1);
            *reinterpret_cast<__nv_bfloat162 *>(lo + o) = __halves2bfloat162(l0, l1);
        }
    }
}

// ------------------------------------------------------------------------------------------------ tc_gemm_kernel
constexpr int G_BM = 128, G_BN = 128, G_STAGES = 3, G_THREADS = 192;
constexpr int G_TILE_BYTES = G_BM * BK * 2;                 // 16 KB per operand tile (BM == BN)
constexpr int G_STAGE_BYTES = 4 * G_TILE_BYTES;             // A_hi, A_lo, W_hi, W_lo
constexpr int G_SMEM = G_STAGES * G_STAGE_BYTES + 1024 + 256;

struct TcGemmParams {
    CUtensorMap a_hi, a_lo, w_hi, w_lo;
    float *C;
    const float *bias;
    long long ldc;
    int M, N, K;
};

// Persistent: CTA i walks output tiles i, i + grid, ... (n fastest, so the n-tiles of one m-tile run side by side and share the A
// tile in L2); two 128-column accumulators in tensor memory, so the epilogue of tile j overlaps the main loop of tile j + 1.
__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + G_STAGES * G_STAGE_BYTES);
    uint64_t *empty = full + G_STAGES;
    uint64_t *acc_full = empty + G_STAGES;   // [2]
    uint64_t *acc_empty = acc_full + 2;      // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (p.K + BK - 1) / BK;
    const int tiles_n = (p.N + G_BN - 1) / G_BN, tiles_m = (p.M + G_BM - 1) / G_BM;
    const int n_tiles = tiles_n * tiles_m;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);  // one arrival per epilogue warp
        }
        mbar_fence_init();
        tma_prefetch_desc(&p.a_hi);
        tma_prefetch_desc(&p.a_lo);
        tma_prefetch_desc(&p.w_hi);
        tma_prefetch_desc(&p.w_lo);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * G_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        {   // ===== TMA producer: the warp walks the loop converged, one elected lane issues (straight UTMALDGs, no ELECT / BRA loop each) =====
            uint32_t g = 0;  // k-chunks issued so far, across tiles: ring stage and phase
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * G_BM, n0 = (tile % tiles_n) * G_BN;
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const uint32_t s = g % G_STAGES;
                    mbar_wait(&empty[s], ((g / G_STAGES) & 1) ^ 1);
                    unsigned char *st = smem + s * G_STAGE_BYTES;
                    if (elect_one_sync()) {
                        mbar_expect_tx(&full[s], G_STAGE_BYTES);
                        tma_load_2d(st + 0 * G_TILE_BYTES, &p.a_hi, &full[s], kc * BK, m0);
                        tma_load_2d(st + 1 * G_TILE_BYTES, &p.a_lo, &full[s], kc * BK, m0);
                        tma_load_2d(st + 2 * G_TILE_BYTES, &p.w_hi, &full[s], kc * BK, n0);
                        tma_load_2d(st + 3 * G_TILE_BYTES, &p.w_lo, &full[s], kc * BK, n0);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        {   // ===== MMA issuer: converged warp, one elected lane (the UTCHMMAs of a k-chunk go out back to back) =====
            constexpr uint32_t idesc = make_idesc_bf16(G_BM, G_BN);
            uint32_t g = 0, t = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
                const uint32_t buf = t & 1;
                mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator (free at first use)
                tc_fence_after();
                const uint32_t acc = tmem_acc + buf * G_BN;
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const uint32_t s = g % G_STAGES;
                    mbar_wait(&full[s], (g / G_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * G_STAGE_BYTES);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                            const uint32_t off = kk * UMMA_K * 2;  // bytes along the 128-byte swizzle row
                            const uint64_t ah = make_sdesc_sw128(st + 0 * G_TILE_BYTES + off), al = make_sdesc_sw128(st + 1 * G_TILE_BYTES + off);
                            const uint64_t wh = make_sdesc_sw128(st + 2 * G_TILE_BYTES + off), wl = make_sdesc_sw128(st + 3 * G_TILE_BYTES + off);
                            umma_bf16(acc, al, wh, idesc, (kc | kk) != 0);
                            umma_bf16(acc, ah, wl, idesc, 1);
                            umma_bf16(acc, ah, wh, idesc, 1);
                        }
                        umma_commit(&empty[s]);
                    }
                    __syncwarp();
                }
                if (elect_one_sync()) umma_commit(&acc_full[buf]);
                __syncwarp();
            }
        }
    } else {  // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        uint32_t t = 0;
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const int m0 = (tile / tiles_n) * G_BM, n0 = (tile % tiles_n) * G_BN;
            const uint32_t buf = t & 1;
            mbar_wait(&acc_full[buf], (t >> 1) & 1);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < G_BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_acc + buf * G_BN + ((uint32_t)(q * 32) << 16) + c * 32, r);
                tmem_ld_wait();
                if (c == G_BN / 32 - 1) {  // the accumulator is in registers: hand it back before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                if (row < p.M) {
                    float *dst = p.C + (size_t)row * p.ldc + n0 + c * 32;
                    const int nb = n0 + c * 32;
                    if (vec_ok && nb + 32 <= p.N) {  // 128-bit stores: a thread owns 128 contiguous bytes of its row
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4 *>(p.bias + nb + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            *reinterpret_cast<float4 *>(dst + j) = make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                                                                               __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = nb + j;
                            if (n < p.N) dst[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n] : 0.f);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, 2 * G_BN);
}

}  // namespace

// C[M][N] (fp32, ldc) = A W^T + bias with device bf16 hi/lo operands, both [rows][K] row-major (K % 8 == 0).
cudaError_t launch_tc_gemm(Ctx *c, const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_lo, const __nv_bfloat16 *w_hi,
                           const __nv_bfloat16 *w_lo, const float *bias, float *C, long long ldc, int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
    TcGemmParams p;
    cudaError_t e;
    if ((e = make_tmap_bf16(&p.a_hi, a_hi, (uint64_t)M, (uint64_t)K, (uint64_t)K, G_BM)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.a_lo, a_lo, (uint64_t)M, (uint64_t)K, (uint64_t)K, G_BM)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.w_hi, w_hi, (uint64_t)N, (uint64_t)K, (uint64_t)K, G_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.w_lo, w_lo, (uint64_t)N, (uint64_t)K, (uint64_t)K, G_BN)) != cudaSuccess) return e;
    p.C = C;
    p.bias = bias;
    p.ldc = ldc;
    p.M = M;
    p.N = N;
    p.K = K;
    if (!c->attr_gemm) {
        if ((e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM)) != cudaSuccess) return e;
        c->attr_gemm = true;
    }
    const long long n_tiles = (long long)((N + G_BN - 1) / G_BN) * ((M + G_BM - 1) / G_BM);
    const int grid = (int)std::max<long long>(1, std::min<long long>(n_tiles, c->sm_count));
    tc_gemm_kernel<<<grid, G_THREADS, G_SMEM, c->stream>>>(p);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_split_rows(Ctx *c, const float *x, size_t ldx, __nv_bfloat16 *hi, __nv_bfloat16 *lo, size_t ldo,
                              size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return cudaSuccess;
    const size_t n = rows * cols;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 16);
    split_rows_kernel<<<blocks, 256, 0, c->stream>>>(x, ldx, hi, lo, ldo, rows, cols);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_split_transpose_enc(Ctx *c, const float *enc, int B, int T, const int *lens_dev, const int *eoff_dev,
                                       int row_base, __nv_bfloat16 *hi, __nv_bfloat16 *lo, const long long *src_off_dev) {
    if (B <= 0 || T <= 0) return cudaSuccess;
    dim3 grid((T + 31) / 32, kEnc / 128, B), block(256);
    split_transpose_enc_kernel<<<grid, block, 0, c->stream>>>(enc, T, lens_dev, eoff_dev, src_off_dev, row_base, hi, lo);
    c->launches++;
    return cudaGetLastError();
}


// ---- host side: split-bf16 weight operands, the hoisted encoder projection, and the launch of the decode kernel ----
constexpr int V_PAD_WS = 17 * 64;  // vocabulary rows padded to whole 64-row slices (decoder_ws.cu)

void decoder_tc_free(TcWeights *w) {
    if (!w) return;
    for (__nv_bfloat16 *p : {w->whh0_hi, w->whh0_lo, w->w1_hi, w->w1_lo, w->wp_hi, w->wp_lo, w->wo_hi, w->wo_lo, w->we_hi, w->we_lo})
        if (p) cudaFree(p);
    delete w;
}

cudaError_t decoder_tc_prepare_weights(Ctx *c) {
    SharedDev *d = c->shared.get();
    if (!d->tc) d->tc = new (std::nothrow) TcWeights();
    if (!d->tc) return cudaErrorMemoryAllocation;
    TcWeights *w = d->tc;
    const BlobLayout L = blob_layout();
    cudaError_t e;
    auto alloc = [&](__nv_bfloat16 **p, size_t n) -> cudaError_t { return *p ? cudaSuccess : cudaMalloc(p, sizeof(__nv_bfloat16) * n); };
    if ((e = alloc(&w->whh0_hi, (size_t)kG * kH)) != cudaSuccess || (e = alloc(&w->whh0_lo, (size_t)kG * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->w1_hi, (size_t)kG * 2 * kH)) != cudaSuccess || (e = alloc(&w->w1_lo, (size_t)kG * 2 * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->wp_hi, (size_t)kH * kH)) != cudaSuccess || (e = alloc(&w->wp_lo, (size_t)kH * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->wo_hi, (size_t)V_PAD_WS * kH)) != cudaSuccess || (e = alloc(&w->wo_lo, (size_t)V_PAD_WS * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->we_hi, (size_t)kH * kEnc)) != cudaSuccess || (e = alloc(&w->we_lo, (size_t)kH * kEnc)) != cudaSuccess) return e;
    cudaMemsetAsync(w->wo_hi, 0, sizeof(__nv_bfloat16) * (size_t)V_PAD_WS * kH, c->stream);
    cudaMemsetAsync(w->wo_lo, 0, sizeof(__nv_bfloat16) * (size_t)V_PAD_WS * kH, c->stream);
    if ((e = launch_split_rows(c, d->whh0p, kH, w->whh0_hi, w->whh0_lo, kH, kG, kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, d->w1p, 2 * kH, w->w1_hi, w->w1_lo, 2 * kH, kG, 2 * kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, c->w_blob + L.w_pred, kH, w->wp_hi, w->wp_lo, kH, kH, kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, c->w_blob + L.w_out, kH, w->wo_hi, w->wo_lo, kH, kV, kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, c->w_blob + L.w_enc, kEnc, w->we_hi, w->we_lo, kEnc, kH, kEnc)) != cudaSuccess) return e;
    return decoder_ws_prepare(c, w);
}

static size_t tc_align(size_t x) { return (x + 1023) & ~(size_t)1023; }

cudaError_t launch_greedy_decode_tc(Ctx *c, const float *enc_dev, const float *enc_host, int B, int T, const int32_t *lens_dev,
                                    const int32_t *lens_host, const int32_t *slots_dev, float *s1_dev, float *s2_dev,
                                    int32_t *tokens_dev, int32_t *ntok_dev, int32_t *nsteps_dev, const int64_t *enc_off_host, int32_t *last_dev) {
    // enc_off_host != nullptr: packed encoder outputs — stream b is a [1024][lens[b]] block at enc + enc_off_host[b]
    DecoderPriv *d = c->dec;
    TcWeights *w = d->tc;
    const int Tq = T > 0 ? T : 1;
    if (!decoder_ws_supported(c)) return cudaErrorNotSupported;  // the weight-stationary kernel needs 147 co-resident CTAs
    cudaError_t e;
    // metadata block (pinned, one upload): rowinfo[B] int4 | src_off[B+1] (packed input) | eoff[B+1]
    const size_t m_ri = 0, m_soff = m_ri + sizeof(int4) * (size_t)B, m_eoff = m_soff + sizeof(long long) * ((size_t)B + 1);
    const size_t meta_bytes = m_eoff + sizeof(int) * ((size_t)B + 1);
    if ((e = c->pin[1].reserve(meta_bytes)) != cudaSuccess) return e;
    char *h_meta = c->pin[1].as<char>();
    long long *h_soff = reinterpret_cast<long long *>(h_meta + m_soff);
    for (int i = 0; i <= B; ++i) h_soff[i] = enc_off_host ? (long long)enc_off_host[i] : 0;
    int *h_eoff = reinterpret_cast<int *>(h_meta + m_eoff);  // first packed row of each stream's valid frames in E
    h_eoff[0] = 0;
    for (int i = 0; i < B; ++i) h_eoff[i + 1] = h_eoff[i] + lens_host[i];
    // the streams, longest first, share the lanes of plan.MT M-tiles: every M-tile lives for the whole kernel
    const WsPlan plan = ws_plan_lanes(lens_host, h_eoff, B, reinterpret_cast<int4 *>(h_meta + m_ri));

    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += tc_align(bytes); return o; };
    const size_t oE = take(sizeof(float) * (size_t)B * Tq * kH);
    const size_t oeh = take(2 * (size_t)B * Tq * kEnc), oel = take(2 * (size_t)B * Tq * kEnc);
    const size_t ometa = take(meta_bytes);
    size_t ws_bytes = 0;
    if ((e = launch_greedy_ws(c, nullptr, B, plan, T, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &ws_bytes)) != cudaSuccess) return e;
    const size_t ows = take(ws_bytes);
    if ((e = d->work.reserve(off)) != cudaSuccess) return e;
    char *base = d->work.as<char>();
    if ((e = cudaMemcpyAsync(base + ometa, h_meta, meta_bytes, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
    const long long *soff_dev = enc_off_host ? reinterpret_cast<const long long *>(base + ometa + m_soff) : nullptr;
    const int *eoff_dev = reinterpret_cast<const int *>(base + ometa + m_eoff);

    float *E = reinterpret_cast<float *>(base + oE);
    if (T > 0) {  // hoisted encoder projection on tcgen05: E[(b,t)][:] = W_enc enc[b][:, t] + b_enc + b_pred
        __nv_bfloat16 *eh = reinterpret_cast<__nv_bfloat16 *>(base + oeh), *el = reinterpret_cast<__nv_bfloat16 *>(base + oel);
        ProfScope prof(c, PK_ENC_PROJ);
        // host-resident encoder outputs: upload chunk k+1 (h2d_stream) while chunk k is split and projected (stream).  At most
        // three uploads are queued at a time (the host waits on the event of chunk k-3): the copy engine is FIFO across
        // streams, and a burst of bulk copies would starve the copies of any other context sharing the GPU.
        const size_t enc_bytes = sizeof(float) * (enc_off_host ? (size_t)enc_off_host[B] : (size_t)B * kEnc * T);
        int by_bytes = (int)std::min<size_t>((size_t)2 * Ctx::kMaxChunks, enc_bytes / ((size_t)24 << 20));  // >= 24 MB per chunk
        if (const char *f = getenv("AMIRA_FORCE_CHUNKS")) by_bytes = std::min(2 * Ctx::kMaxChunks, atoi(f));  // tests: chunked path at small sizes
        const int n_chunks = enc_host ? std::max(1, std::min(by_bytes, B / 32)) : 1;
        // Batch-sized uploads of the lanes of one GPU take turns, first come first served, instead of sharing the link chunk by
        // chunk: with several batches in flight the call that came first gets its encoder outputs — and with them its decode
        // kernel — first, and the SMs do not wait for an upload that is one of four advancing at a quarter of the rate (e2e leg of
        // bench.py, four steps in flight, A/B on one box: 33.9 -> 32.1 ms per step, and the same from run to run).
        // AMIRA_H2D_FIFO=0 restores the free-for-all.
        static const int fifo_env = upload_fifo_mode();
        std::unique_lock<std::mutex> turn;
        if (fifo_env && enc_host && n_chunks > 3) turn = std::unique_lock<std::mutex>(upload_turn_mutex(c->device));
        for (int k = 0; k < n_chunks; ++k) {
            const int b0 = (int)((long long)B * k / n_chunks), b1 = (int)((long long)B * (k + 1) / n_chunks);
            if (b1 <= b0) continue;
            const size_t eo = enc_off_host ? (size_t)enc_off_host[b0] : (size_t)b0 * kEnc * T, ro = (size_t)h_eoff[b0];
            const size_t en = (enc_off_host ? (size_t)enc_off_host[b1] : (size_t)b1 * kEnc * T) - eo;
            const int rows = h_eoff[b1] - h_eoff[b0];
            if (enc_host) {
                if (k >= 3 && (e = cudaEventSynchronize(c->ev_pool[2 * Ctx::kMaxChunks + k - 3])) != cudaSuccess) return e;
                if (en > 0 && (e = cudaMemcpyAsync(const_cast<float *>(enc_dev) + eo, enc_host + eo, sizeof(float) * en, cudaMemcpyHostToDevice,
                                                   c->h2d_stream)) != cudaSuccess) return e;
                if ((e = cudaEventRecord(c->ev_pool[2 * Ctx::kMaxChunks + k], c->h2d_stream)) != cudaSuccess) return e;
                if ((e = cudaStreamWaitEvent(c->stream, c->ev_pool[2 * Ctx::kMaxChunks + k], 0)) != cudaSuccess) return e;
            }
            if ((e = launch_split_transpose_enc(c, enc_off_host ? enc_dev : enc_dev + eo, b1 - b0, T, lens_dev + b0, eoff_dev + b0, h_eoff[b0],
                                                eh + ro * kEnc, el + ro * kEnc, soff_dev ? soff_dev + b0 : nullptr)) != cudaSuccess) return e;
            if ((e = launch_tc_gemm(c, eh + ro * kEnc, el + ro * kEnc, w->we_hi, w->we_lo, d->bjoint, E + ro * kH, kH, rows, kH, kEnc)) != cudaSuccess) return e;
        }
    }
    return launch_greedy_ws(c, E, B, plan, T, reinterpret_cast<const int4 *>(base + ometa + m_ri), slots_dev, s1_dev, s2_dev, tokens_dev,
                            ntok_dev, nsteps_dev, base + ows, &ws_bytes, last_dev);
}

}  // namespace amira

// ---- diagnostics entry (tests/test_gpu_tcgen05.py): C = A W^T + bias through the tcgen05 split-bf16 path ----
extern "C" int32_t amira_debug_tc_gemm(amira_ctx *ctx, const float *A, const float *W, const float *bias, int32_t M, int32_t N,
                                       int32_t K, float *C) {
    using namespace amira;
    if (!ctx || !A || !W || !C || M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0) return AMIRA_ERR_INVALID_VALUE;
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    std::lock_guard<std::mutex> lock(c->mu);
    cudaSetDevice(c->device);
    float *dA = nullptr, *dW = nullptr, *dC = nullptr, *dB = nullptr;
    __nv_bfloat16 *ah = nullptr, *al = nullptr, *wh = nullptr, *wl = nullptr;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; return e == cudaSuccess; };
    ok(cudaMalloc(&dA, sizeof(float) * (size_t)M * K));
    ok(cudaMalloc(&dW, sizeof(float) * (size_t)N * K));
    ok(cudaMalloc(&dC, sizeof(float) * (size_t)M * N));
    ok(cudaMalloc(&dB, sizeof(float) * (size_t)N));
    ok(cudaMalloc(&ah, 2 * (size_t)M * K));
    ok(cudaMalloc(&al, 2 * (size_t)M * K));
    ok(cudaMalloc(&wh, 2 * (size_t)N * K));
    ok(cudaMalloc(&wl, 2 * (size_t)N * K));
    if (e == cudaSuccess) {
        ok(cudaMemcpyAsync(dA, A, sizeof(float) * (size_t)M * K, cudaMemcpyHostToDevice, c->stream));
        ok(cudaMemcpyAsync(dW, W, sizeof(float) * (size_t)N * K, cudaMemcpyHostToDevice, c->stream));
        if (bias) ok(cudaMemcpyAsync(dB, bias, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, c->stream));
        ok(launch_split_rows(c, dA, K, ah, al, K, M, K));
        ok(launch_split_rows(c, dW, K, wh, wl, K, N, K));
        ok(launch_tc_gemm(c, ah, al, wh, wl, bias ? dB : nullptr, dC, N, M, N, K));
        ok(cudaMemcpyAsync(C, dC, sizeof(float) * (size_t)M * N, cudaMemcpyDeviceToHost, c->stream));
        ok(cudaStreamSynchronize(c->stream));
    }
    for (void *p : {(void *)dA, (void *)dW, (void *)dC, (void *)dB, (void *)ah, (void *)al, (void *)wh, (void *)wl})
        if (p) cudaFree(p);
    if (e != cudaSuccess) {
        c->err = std::string("amira_debug_tc_gemm: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return AMIRA_ERR_UNKNOWN;
    }
    return AMIRA_OK;
}
