// decoder_tc.cu — tcgen05 (5th-gen tensor core) paths of the decoder_joint stage on sm_100a.
//
//   tc_gemm_kernel        C[M][N] = A[M][K] W[N][K]^T (+bias) with split-bf16 operands (hi*hi + lo*hi + hi*lo, fp32
//                         accumulate in TMEM): TMA -> 128B-swizzled smem ring -> tcgen05.mma issued by one thread ->
//                         tcgen05.ld epilogue.  Used for the hoisted encoder projection E = enc W_enc^T + b
//                         (M = sum of encoder frames, the one large GEMM of the path) and as the unit-testable proof of
//                         the descriptor / swizzle conventions (tests/test_gpu_tcgen05.py).
//   greedy_tc_kernel      the persistent greedy-decode loop with the four per-iteration GEMMs on tcgen05
//                         (decode_engine = 2); same control flow and algebra as decoder.cu.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

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
        if (lane == 0) {  // ===== TMA producer =====
            uint32_t g = 0;  // k-chunks issued so far, across tiles: ring stage and phase
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * G_BM, n0 = (tile % tiles_n) * G_BN;
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const uint32_t s = g % G_STAGES;
                    mbar_wait(&empty[s], ((g / G_STAGES) & 1) ^ 1);
                    unsigned char *st = smem + s * G_STAGE_BYTES;
                    mbar_expect_tx(&full[s], G_STAGE_BYTES);
                    tma_load_2d(st + 0 * G_TILE_BYTES, &p.a_hi, &full[s], kc * BK, m0);
                    tma_load_2d(st + 1 * G_TILE_BYTES, &p.a_lo, &full[s], kc * BK, m0);
                    tma_load_2d(st + 2 * G_TILE_BYTES, &p.w_hi, &full[s], kc * BK, n0);
                    tma_load_2d(st + 3 * G_TILE_BYTES, &p.w_lo, &full[s], kc * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
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
                umma_commit(&acc_full[buf]);
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


// =================================================================================================================
// greedy_tc_kernel — the persistent greedy-decode loop on tcgen05 (decode_engine = 2).
//
// Same algebra and control flow as decoder.cu (citations there); differences:
//  * rows (streams) are sorted by encoded length so active rows stay a prefix and whole 128-row M-tiles retire;
//  * activations live in HBM/L2 as split bf16 (hi, lo) [parity][Mpad][640] next to an fp32 copy for the final state;
//  * each phase is a set of 128x128 output tiles; warp 0 = TMA producer, warp 1 = MMA issuer (three tcgen05.mma per
//    k-step: lo*hi + hi*lo + hi*hi), warps 2-5 = epilogue (one thread per stream row reads its TMEM lane and applies
//    the LSTM cell / tanh / bias+argmax); TMEM holds two accumulators so the next tile's MMAs overlap the epilogue;
//  * the per-stream control update runs in the CTA that finishes the last vocab tile of an M-tile (atomic ticket), so
//    an iteration needs four grid syncs and no redundant work.
namespace {

constexpr int T_BM = 128, T_BN = 128, T_STAGES = 3, T_THREADS = 192;
constexpr int T_TILE_BYTES = T_BM * BK * 2;
constexpr int T_STAGE_BYTES = 4 * T_TILE_BYTES;
constexpr int NT_G = kG / T_BN;                       // 20 gate tiles
constexpr int NT_P = kH / T_BN;                       // 5
constexpr int NT_O = (kV + T_BN - 1) / T_BN;          // 9
constexpr int V_PAD_TC = NT_O * T_BN;                 // 1152
constexpr int MAX_MT = 256;
constexpr int KC = kH / BK;                           // 10 k-chunks per 640-wide segment
constexpr int T_SMEM = T_STAGES * T_STAGE_BYTES + 1024 + 2048;

struct TCtl {
    int t, sym, total, last, active, nsteps, failed, pad;
};

struct TcDecParams {
    CUtensorMap h0_hi, h0_lo, h1_hi, h1_lo, z_hi, z_lo;
    CUtensorMap whh0_hi, whh0_lo, w1_hi, w1_lo, wp_hi, wp_lo, wo_hi, wo_lo;
    const float *g0p, *b1p, *boutp, *E;
    int B, Mpad, MT, T;
    const int *lens, *slots, *perm, *eoff;
    __nv_bfloat16 *h0b_hi, *h0b_lo, *h1b_hi, *h1b_lo, *zb_hi, *zb_lo;
    float *h0f, *h1f, *c0, *c1;
    float *pval;
    int *pidx;
    TCtl *ctl;
    int *tile_active, *done_cnt, *total_active;  // total_active: [0..1] by parity, [2] failed streams
    int *cnt_a, *cnt_b, *cnt_c, *ctl_done, *dead_at;  // dataflow engine: per-M-tile monotonic counters
    float *s1, *s2;
    int *tokens, *ntok, *nsteps;
    int max_sym, max_total, blank, relu;
};

struct TcSmemCtl {
    uint64_t full[T_STAGES], empty[T_STAGES], acc_full[2], acc_empty[2];
    uint32_t tmem_slot;
    int total, n_mt, last_flag, act_cnt;
    int tile_active[MAX_MT];
};

struct PipeState {
    uint32_t k;     // k-chunk counter (producer / MMA)
    uint32_t tile;  // processed-tile counter (MMA / epilogue)
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ TCtl load_ctl(const TCtl *q) {  // L1-bypassing: written by another SM's control update
    const int4 a = __ldcg(reinterpret_cast<const int4 *>(q)), b = __ldcg(reinterpret_cast<const int4 *>(q) + 1);
    TCtl c;
    c.t = a.x; c.sym = a.y; c.total = a.z; c.last = a.w; c.active = b.x; c.nsteps = b.y; c.failed = b.z; c.pad = b.w;
    return c;
}

__device__ __forceinline__ size_t tc_state_off(const TcDecParams &p, int layer, int b) {
    return p.slots ? ((size_t)p.slots[b] * 2 + layer) * kH : ((size_t)layer * p.B + b) * kH;
}

enum { PH_A = 0, PH_B = 1, PH_C = 2, PH_D = 3 };

template <int PH>
__device__ __forceinline__ void run_phase(const TcDecParams &p, unsigned char *stages, TcSmemCtl &sc, int it, PipeState &ps) {
    constexpr int NT = (PH == PH_A || PH == PH_B) ? NT_G : (PH == PH_C ? NT_P : NT_O);
    constexpr int NSEG = PH == PH_B ? 2 : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int par = it & 1;
    const int ntiles = sc.n_mt * NT;
    const uint32_t tmem_base = sc.tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===================== TMA producer =====================
            fence_proxy_async();
            const CUtensorMap *a_hi[2], *a_lo[2], *w_hi, *w_lo;
            int a_row[2];
            if (PH == PH_A) { a_hi[0] = &p.h0_hi; a_lo[0] = &p.h0_lo; a_row[0] = par * p.Mpad; w_hi = &p.whh0_hi; w_lo = &p.whh0_lo; }
            else if (PH == PH_B) {
                a_hi[0] = &p.h0_hi; a_lo[0] = &p.h0_lo; a_row[0] = (par ^ 1) * p.Mpad;
                a_hi[1] = &p.h1_hi; a_lo[1] = &p.h1_lo; a_row[1] = par * p.Mpad;
                w_hi = &p.w1_hi; w_lo = &p.w1_lo;
            } else if (PH == PH_C) { a_hi[0] = &p.h1_hi; a_lo[0] = &p.h1_lo; a_row[0] = (par ^ 1) * p.Mpad; w_hi = &p.wp_hi; w_lo = &p.wp_lo; }
            else { a_hi[0] = &p.z_hi; a_lo[0] = &p.z_lo; a_row[0] = 0; w_hi = &p.wo_hi; w_lo = &p.wo_lo; }
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int mt = tile / NT, nt = tile % NT;
                if (sc.tile_active[mt] == 0) continue;
#pragma unroll 1
                for (int seg = 0; seg < NSEG; ++seg)
#pragma unroll 1
                    for (int kc = 0; kc < KC; ++kc) {
                        const uint32_t s = ps.k % T_STAGES;
                        mbar_wait(&sc.empty[s], ((ps.k / T_STAGES) & 1) ^ 1);
                        unsigned char *st = stages + s * T_STAGE_BYTES;
                        mbar_expect_tx(&sc.full[s], T_STAGE_BYTES);
                        tma_load_2d(st + 0 * T_TILE_BYTES, a_hi[seg], &sc.full[s], kc * BK, a_row[seg] + mt * T_BM);
                        tma_load_2d(st + 1 * T_TILE_BYTES, a_lo[seg], &sc.full[s], kc * BK, a_row[seg] + mt * T_BM);
                        tma_load_2d(st + 2 * T_TILE_BYTES, w_hi, &sc.full[s], seg * kH + kc * BK, nt * T_BN);
                        tma_load_2d(st + 3 * T_TILE_BYTES, w_lo, &sc.full[s], seg * kH + kc * BK, nt * T_BN);
                        ++ps.k;
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc_bf16(T_BM, T_BN);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int mt = tile / NT;
                if (sc.tile_active[mt] == 0) continue;
                const uint32_t buf = ps.tile & 1, use = ps.tile >> 1;
                mbar_wait(&sc.acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * T_BN;
#pragma unroll 1
                for (int kc = 0; kc < NSEG * KC; ++kc) {
                    const uint32_t s = ps.k % T_STAGES;
                    mbar_wait(&sc.full[s], (ps.k / T_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t st = smem_u32(stages + s * T_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                        const uint32_t off = kk * UMMA_K * 2;
                        const uint64_t ah = make_sdesc_sw128(st + 0 * T_TILE_BYTES + off), al = make_sdesc_sw128(st + 1 * T_TILE_BYTES + off);
                        const uint64_t wh = make_sdesc_sw128(st + 2 * T_TILE_BYTES + off), wl = make_sdesc_sw128(st + 3 * T_TILE_BYTES + off);
                        umma_bf16(acc, al, wh, idesc, (kc | kk) != 0);
                        umma_bf16(acc, ah, wl, idesc, 1);
                        umma_bf16(acc, ah, wh, idesc, 1);
                    }
                    umma_commit(&sc.empty[s]);
                    ++ps.k;
                }
                umma_commit(&sc.acc_full[buf]);
                ++ps.tile;
            }
        }
    } else {  // ===================== epilogue: one thread per stream row =====================
        const int q = warp & 3, etid = q * 32 + lane;  // TMEM lane == row within the M-tile
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int mt = tile / NT, nt = tile % NT;
            if (sc.tile_active[mt] == 0) continue;
            const uint32_t buf = ps.tile & 1, use = ps.tile >> 1;
            const int row = mt * T_BM + etid, n0 = nt * T_BN;
            const TCtl c = load_ctl(p.ctl + row);
            mbar_wait(&sc.acc_full[buf], use & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * T_BN + ((uint32_t)(q * 32) << 16);
            float best_v = -INFINITY;
            int best_i = 0x7fffffff;
#pragma unroll 1
            for (int cc = 0; cc < T_BN / 32; ++cc) {
                uint32_t r[32];
                tmem_ld32(acc + cc * 32, r);
                tmem_ld_wait();
                if (!c.active) continue;
                const int nb = n0 + cc * 32;
                if (PH == PH_A || PH == PH_B) {
                    const float *addp = PH == PH_A ? p.g0p + (size_t)c.last * kG + nb : p.b1p + nb;
                    float *cst = (PH == PH_A ? p.c0 : p.c1) + (size_t)row * kH + nb / 4;
                    float *hf = (PH == PH_A ? p.h0f : p.h1f) + (size_t)row * kH + nb / 4;
                    const size_t ob = ((size_t)(par ^ 1) * p.Mpad + row) * kH + nb / 4;
                    __nv_bfloat16 *bh = (PH == PH_A ? p.h0b_hi : p.h1b_hi) + ob, *bl = (PH == PH_A ? p.h0b_lo : p.h1b_lo) + ob;
                    float cold[8], hnew[8];
                    *reinterpret_cast<float4 *>(&cold[0]) = __ldcg(reinterpret_cast<const float4 *>(cst));
                    *reinterpret_cast<float4 *>(&cold[4]) = __ldcg(reinterpret_cast<const float4 *>(cst) + 1);
                    __align__(16) __nv_bfloat16 vh[8], vl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 ad = __ldg(reinterpret_cast<const float4 *>(addp) + j);
                        const float gi = sigm(__uint_as_float(r[4 * j + 0]) + ad.x), gf = sigm(__uint_as_float(r[4 * j + 1]) + ad.y);
                        const float gg = tanhf(__uint_as_float(r[4 * j + 2]) + ad.z), go = sigm(__uint_as_float(r[4 * j + 3]) + ad.w);
                        const float cn = gf * cold[j] + gi * gg;
                        cold[j] = cn;
                        hnew[j] = go * tanhf(cn);
                        split_bf16(hnew[j], vh[j], vl[j]);
                    }
                    reinterpret_cast<float4 *>(cst)[0] = *reinterpret_cast<float4 *>(&cold[0]);
                    reinterpret_cast<float4 *>(cst)[1] = *reinterpret_cast<float4 *>(&cold[4]);
                    reinterpret_cast<float4 *>(hf)[0] = *reinterpret_cast<float4 *>(&hnew[0]);
                    reinterpret_cast<float4 *>(hf)[1] = *reinterpret_cast<float4 *>(&hnew[4]);
                    *reinterpret_cast<uint4 *>(bh) = *reinterpret_cast<uint4 *>(vh);
                    *reinterpret_cast<uint4 *>(bl) = *reinterpret_cast<uint4 *>(vl);
                } else if (PH == PH_C) {
                    const float *e = p.E + ((size_t)p.eoff[p.perm[row]] + c.t) * kH + nb;
                    __nv_bfloat16 *bh = p.zb_hi + (size_t)row * kH + nb, *bl = p.zb_lo + (size_t)row * kH + nb;
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        __align__(16) __nv_bfloat16 vh[8], vl[8];
                        const float4 e0 = __ldg(reinterpret_cast<const float4 *>(e) + 2 * j8), e1 = __ldg(reinterpret_cast<const float4 *>(e) + 2 * j8 + 1);
                        const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v = __uint_as_float(r[8 * j8 + j]) + ev[j];
                            split_bf16(p.relu ? fmaxf(v, 0.f) : tanhf(v), vh[j], vl[j]);
                        }
                        reinterpret_cast<uint4 *>(bh)[j8] = *reinterpret_cast<uint4 *>(vh);
                        reinterpret_cast<uint4 *>(bl)[j8] = *reinterpret_cast<uint4 *>(vl);
                    }
                } else {  // PH_D: first-max argmax over this tile's columns (zero_copy.rs:190-232 tie rule)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nb + j;
                        if (n < kV) {
                            const float v = __uint_as_float(r[j]) + __ldg(p.boutp + n);
                            if (v > best_v || best_i == 0x7fffffff) { best_v = v; best_i = n; }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&sc.acc_empty[buf]);
            ++ps.tile;

            if (PH == PH_D) {
                if (c.active) {
                    p.pval[(size_t)row * NT_O + nt] = best_v;
                    p.pidx[(size_t)row * NT_O + nt] = best_i;
                }
                __threadfence();
                named_bar_sync(1, 128);
                if (etid == 0) {
                    const int old = atomicAdd(&p.done_cnt[mt], 1);
                    sc.last_flag = (old == NT_O - 1);
                    sc.act_cnt = 0;
                }
                named_bar_sync(1, 128);
                if (sc.last_flag) {  // this CTA finished the M-tile's last vocab tile: per-stream control update
                    __threadfence();
                    TCtl n = c;
                    if (n.active) {
                        float bv = __ldcg(p.pval + (size_t)row * NT_O);
                        int bi = __ldcg(p.pidx + (size_t)row * NT_O);
                        for (int qn = 1; qn < NT_O; ++qn) {
                            const float v = __ldcg(p.pval + (size_t)row * NT_O + qn);
                            if (v > bv) { bv = v; bi = __ldcg(p.pidx + (size_t)row * NT_O + qn); }
                        }
                        const int len = p.lens[p.perm[row]];
                        n.nsteps += 1;                       // state carried unconditionally (decoder_optimized.rs:154)
                        n.sym += 1;                          // :133
                        if (bi == p.blank) {                 // :171-173
                            n.t += 1; n.sym = 0;
                            if (n.t >= len) n.active = 0;
                        } else {
                            p.tokens[(size_t)p.perm[row] * p.max_total + n.total] = bi;   // :176
                            n.total += 1;
                            n.last = bi;
                            if (n.total >= p.max_total) n.active = 0;                    // :179-188
                            else if (n.sym >= p.max_sym) {                               // :133-137
                                n.t += 1; n.sym = 0;
                                if (n.t >= len) n.active = 0;
                            }
                            if (n.active && bi >= kEmbRows) { n.active = 0; n.failed = 1; }  // next step would fail (:148-152)
                        }
                        p.ctl[row] = n;
                        if (n.active) atomicAdd(&sc.act_cnt, 1);
                    }
                    named_bar_sync(1, 128);
                    if (etid == 0) {
                        p.tile_active[mt] = sc.act_cnt;
                        p.done_cnt[mt] = 0;
                        if (sc.act_cnt) atomicAdd(&p.total_active[par ^ 1], sc.act_cnt);
                    }
                }
                named_bar_sync(1, 128);
            }
        }
        __threadfence();
        fence_proxy_async();
    }
}

__global__ void __launch_bounds__(T_THREADS, 1) greedy_tc_kernel(const __grid_constant__ TcDecParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *stages = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    TcSmemCtl &sc = *reinterpret_cast<TcSmemCtl *>(stages + T_STAGES * T_STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            mbar_init(&sc.full[s], 1);
            mbar_init(&sc.empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sc.acc_full[b], 1);
            mbar_init(&sc.acc_empty[b], 128);
        }
        mbar_fence_init();
        const CUtensorMap *maps[14] = {&p.h0_hi, &p.h0_lo, &p.h1_hi, &p.h1_lo, &p.z_hi, &p.z_lo, &p.whh0_hi, &p.whh0_lo,
                                       &p.w1_hi, &p.w1_lo, &p.wp_hi, &p.wp_lo, &p.wo_hi, &p.wo_lo};
        for (int i = 0; i < 14; ++i) tma_prefetch_desc(maps[i]);
    }
    if (warp == 1) tmem_alloc(&sc.tmem_slot, 2 * T_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- prologue: initial LSTM state (fp32 + split bf16, parity 0), control, activity counters ----
    const size_t n_state = (size_t)p.Mpad * kH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n_state; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / kH), j = (int)(i % kH);
        const int b = row < p.B ? p.perm[row] : -1;
        const float h0 = (b >= 0 && p.s1) ? p.s1[tc_state_off(p, 0, b) + j] : 0.f;
        const float h1 = (b >= 0 && p.s1) ? p.s1[tc_state_off(p, 1, b) + j] : 0.f;
        p.h0f[i] = h0;
        p.h1f[i] = h1;
        p.c0[i] = (b >= 0 && p.s2) ? p.s2[tc_state_off(p, 0, b) + j] : 0.f;
        p.c1[i] = (b >= 0 && p.s2) ? p.s2[tc_state_off(p, 1, b) + j] : 0.f;
        __nv_bfloat16 hh, hl;
        split_bf16(h0, hh, hl);
        p.h0b_hi[i] = hh; p.h0b_lo[i] = hl;
        split_bf16(h1, hh, hl);
        p.h1b_hi[i] = hh; p.h1b_lo[i] = hl;
    }
    for (int row = blockIdx.x * blockDim.x + tid; row < p.Mpad; row += gridDim.x * blockDim.x) {
        TCtl c;
        c.t = 0; c.sym = 0; c.total = 0; c.last = p.blank; c.nsteps = 0; c.failed = 0; c.pad = 0;
        c.active = (row < p.B && p.lens[p.perm[row]] > 0) ? 1 : 0;
        p.ctl[row] = c;
        if (c.active) {
            atomicAdd(&p.tile_active[row / T_BM], 1);
            atomicAdd(&p.total_active[0], 1);
        }
    }
    __threadfence();
    fence_proxy_async();
    grid.sync();

    PipeState ps{0, 0};
    int it = 0;
    for (;; ++it) {
        const int par = it & 1;
        __syncthreads();
        if (tid == 0) sc.total = __ldcg(p.total_active + par);
        for (int i = tid; i < p.MT; i += T_THREADS) sc.tile_active[i] = __ldcg(p.tile_active + i);
        __syncthreads();
        if (tid == 0) {
            int n = 0;
            for (int i = 0; i < p.MT; ++i)
                if (sc.tile_active[i] > 0) n = i + 1;
            sc.n_mt = n;
        }
        __syncthreads();
        if (sc.total == 0) break;
        if (blockIdx.x == 0 && tid == 0) p.total_active[par ^ 1] = 0;
        run_phase<PH_A>(p, stages, sc, it, ps);
        grid.sync();
        run_phase<PH_B>(p, stages, sc, it, ps);
        grid.sync();
        run_phase<PH_C>(p, stages, sc, it, ps);
        grid.sync();
        run_phase<PH_D>(p, stages, sc, it, ps);
        grid.sync();
    }

    // ---- results ----
    for (int row = blockIdx.x * blockDim.x + tid; row < p.B; row += gridDim.x * blockDim.x) {
        const TCtl c = load_ctl(p.ctl + row);
        const int b = p.perm[row];
        p.ntok[b] = c.failed ? -1 : c.total;
        if (p.nsteps) p.nsteps[b] = c.nsteps;
        if (c.failed) atomicAdd(&p.total_active[2], 1);
    }
    if (p.s1 && p.s2) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < (size_t)p.B * kH; i += (size_t)gridDim.x * blockDim.x) {
            const int row = (int)(i / kH), j = (int)(i % kH), b = p.perm[row];
            p.s1[tc_state_off(p, 0, b) + j] = p.h0f[i];
            p.s1[tc_state_off(p, 1, b) + j] = p.h1f[i];
            p.s2[tc_state_off(p, 0, b) + j] = p.c0[i];
            p.s2[tc_state_off(p, 1, b) + j] = p.c1[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(sc.tmem_slot, 2 * T_BN);
}


// =================================================================================================================
// greedy_df_kernel — dataflow variant (decode_engine = 3, the default).  Same tiles, operands and epilogue math as
// greedy_tc_kernel, but no grid-wide barrier inside the loop:
//  * every (M-tile, phase, N-tile) has a fixed owner CTA for the whole kernel; a CTA walks its tiles in the global order
//    (iteration, phase, M-tile, N-tile), which is consistent with every dependency, so the schedule cannot deadlock;
//  * dependencies are per-M-tile monotonic counters in global memory (L0 tiles done -> L1 may start, ...), released
//    with __threadfence + atomicAdd by the epilogue and acquired by the scheduler thread before it issues the TMA loads;
//    M-tiles therefore advance independently and one tile's epilogue overlaps other tiles' TMA and MMA work;
//  * warp 0 lane 0 is scheduler + TMA producer and publishes tile descriptors to the MMA warp and the 16 epilogue warps
//    through a small mbarrier-guarded queue in shared memory;
//  * the epilogue is spread over 16 warps: four per TMEM lane quarter, 32 accumulator columns (8 hidden units) each;
//  * the CTA that finishes an M-tile's last vocab tile runs the per-stream control update, and when the M-tile has no
//    active stream left it writes that M-tile's results and marks it dead; CTAs exit when all their M-tiles are dead.
// Spin loops carry a cycle-count watchdog that traps instead of hanging the GPU.
constexpr int D_EPI_WARPS = 16, D_THREADS = (4 + D_EPI_WARPS) * 32;  // 640
constexpr int D_EPI_THREADS = D_EPI_WARPS * 32;                       // 512
constexpr int D_Q = 4;
constexpr int D_TILES_PER_MT = 2 * NT_G + NT_P + NT_O;                // 54
constexpr int D_MAX_LIST = 192;
constexpr long long D_SPIN_LIMIT = 6000000000LL;                      // ~3 s of SM clocks
constexpr int D_SMEM = T_STAGES * T_STAGE_BYTES + 1024 + 4096;
constexpr int NPART = NT_O * 4;                                       // argmax partials per row

struct DfDesc {
    int phase, mt, nt, it;
};
struct DfSmem {
    uint64_t full[T_STAGES], empty[T_STAGES], acc_full[2], acc_empty[2], q_full[D_Q], q_empty[D_Q];
    uint32_t tmem_slot;
    int last_flag, act_cnt, n_list;
    DfDesc q[D_Q];
    unsigned char list_phase[D_MAX_LIST], list_mt[D_MAX_LIST], list_nt[D_MAX_LIST], dead[MAX_MT];
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until_ge(const int *p, int target) {
    if (ld_acquire(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire(p) < target) {
        __nanosleep(32);
        if (clock64() - t0 > D_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++n & 0xfff) == 0 && clock64() - t0 > D_SPIN_LIMIT) __trap();
    }
}

__global__ void __launch_bounds__(D_THREADS, 1) greedy_df_kernel(const __grid_constant__ TcDecParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *stages = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    DfSmem &sm = *reinterpret_cast<DfSmem *>(stages + T_STAGES * T_STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.acc_full[b], 1);
            mbar_init(&sm.acc_empty[b], D_EPI_THREADS);
        }
        for (int i = 0; i < D_Q; ++i) {
            mbar_init(&sm.q_full[i], 1);
            mbar_init(&sm.q_empty[i], 1 + D_EPI_WARPS);
        }
        mbar_fence_init();
        const CUtensorMap *maps[14] = {&p.h0_hi, &p.h0_lo, &p.h1_hi, &p.h1_lo, &p.z_hi, &p.z_lo, &p.whh0_hi, &p.whh0_lo,
                                       &p.w1_hi, &p.w1_lo, &p.wp_hi, &p.wp_lo, &p.wo_hi, &p.wo_lo};
        for (int i = 0; i < 14; ++i) tma_prefetch_desc(maps[i]);
        // this CTA's tiles, in (phase, M-tile, N-tile) order: owner(mt, q) = (mt * 54 + q) mod gridDim
        int n = 0;
        for (int ph = 0; ph < 4; ++ph) {
            const int nt_n = ph < 2 ? NT_G : (ph == 2 ? NT_P : NT_O);
            const int qbase = ph == 0 ? 0 : (ph == 1 ? NT_G : (ph == 2 ? 2 * NT_G : 2 * NT_G + NT_P));
            for (int mt = 0; mt < p.MT; ++mt)
                for (int nt = 0; nt < nt_n; ++nt)
                    if ((mt * D_TILES_PER_MT + qbase + nt) % (int)gridDim.x == (int)blockIdx.x && n < D_MAX_LIST) {
                        sm.list_phase[n] = (unsigned char)ph;
                        sm.list_mt[n] = (unsigned char)mt;
                        sm.list_nt[n] = (unsigned char)nt;
                        ++n;
                    }
        }
        sm.n_list = n;
    }
    for (int i = tid; i < MAX_MT; i += D_THREADS) sm.dead[i] = 0;
    if (warp == 1) tmem_alloc(&sm.tmem_slot, 2 * T_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- prologue: initial LSTM state (fp32 + split bf16, parity 0), control, default results ----
    const size_t n_state = (size_t)p.Mpad * kH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n_state; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / kH), j = (int)(i % kH);
        const int b = row < p.B ? p.perm[row] : -1;
        const float h0 = (b >= 0 && p.s1) ? p.s1[tc_state_off(p, 0, b) + j] : 0.f;
        const float h1 = (b >= 0 && p.s1) ? p.s1[tc_state_off(p, 1, b) + j] : 0.f;
        p.h0f[i] = h0;
        p.h1f[i] = h1;
        p.c0[i] = (b >= 0 && p.s2) ? p.s2[tc_state_off(p, 0, b) + j] : 0.f;
        p.c1[i] = (b >= 0 && p.s2) ? p.s2[tc_state_off(p, 1, b) + j] : 0.f;
        __nv_bfloat16 hh, hl;
        split_bf16(h0, hh, hl);
        p.h0b_hi[i] = hh; p.h0b_lo[i] = hl;
        split_bf16(h1, hh, hl);
        p.h1b_hi[i] = hh; p.h1b_lo[i] = hl;
    }
    for (int row = blockIdx.x * blockDim.x + tid; row < p.Mpad; row += gridDim.x * blockDim.x) {
        TCtl c;
        c.t = 0; c.sym = 0; c.total = 0; c.last = p.blank; c.nsteps = 0; c.failed = 0; c.pad = 0;
        c.active = (row < p.B && p.lens[p.perm[row]] > 0) ? 1 : 0;
        p.ctl[row] = c;
        if (c.active) atomicAdd(&p.tile_active[row / T_BM], 1);
        if (row < p.B) {
            p.ntok[p.perm[row]] = 0;
            if (p.nsteps) p.nsteps[p.perm[row]] = 0;
        }
    }
    __threadfence();
    fence_proxy_async();
    grid.sync();
    // M-tiles with no active stream never start (dead_at = 0); host zero-initialised the counters
    for (int mt = blockIdx.x * blockDim.x + tid; mt < p.MT; mt += gridDim.x * blockDim.x)
        p.dead_at[mt] = __ldcg(p.tile_active + mt) > 0 ? 0x7fffffff : 0;
    __threadfence();
    grid.sync();

    const uint32_t tmem_base = sm.tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===================== scheduler + TMA producer =====================
            uint32_t k = 0, qn = 0;
            for (int it = 0;; ++it) {
                const int par = it & 1;
                bool any = false;
                for (int li = 0; li < sm.n_list; ++li) {
                    const int ph = sm.list_phase[li], mt = sm.list_mt[li], nt = sm.list_nt[li];
                    if (sm.dead[mt]) continue;
                    if (it > 0) spin_until_ge(p.ctl_done + mt, it);       // control of iteration it-1 finished
                    if (ld_acquire(p.dead_at + mt) <= it) { sm.dead[mt] = 1; continue; }
                    any = true;
                    const CUtensorMap *a_hi[2], *a_lo[2], *w_hi, *w_lo;
                    int a_row[2], nseg = 1;
                    if (ph == PH_A) { a_hi[0] = &p.h0_hi; a_lo[0] = &p.h0_lo; a_row[0] = par * p.Mpad; w_hi = &p.whh0_hi; w_lo = &p.whh0_lo; }
                    else if (ph == PH_B) {
                        spin_until_ge(p.cnt_a + mt, NT_G * (it + 1));
                        a_hi[0] = &p.h0_hi; a_lo[0] = &p.h0_lo; a_row[0] = (par ^ 1) * p.Mpad;
                        a_hi[1] = &p.h1_hi; a_lo[1] = &p.h1_lo; a_row[1] = par * p.Mpad;
                        w_hi = &p.w1_hi; w_lo = &p.w1_lo; nseg = 2;
                    } else if (ph == PH_C) {
                        spin_until_ge(p.cnt_b + mt, NT_G * (it + 1));
                        a_hi[0] = &p.h1_hi; a_lo[0] = &p.h1_lo; a_row[0] = (par ^ 1) * p.Mpad; w_hi = &p.wp_hi; w_lo = &p.wp_lo;
                    } else {
                        spin_until_ge(p.cnt_c + mt, NT_P * (it + 1));
                        a_hi[0] = &p.z_hi; a_lo[0] = &p.z_lo; a_row[0] = 0; w_hi = &p.wo_hi; w_lo = &p.wo_lo;
                    }
                    fence_proxy_async();
                    {   // publish the tile
                        const uint32_t slot = qn % D_Q;
                        mbar_wait_wd(&sm.q_empty[slot], ((qn / D_Q) & 1) ^ 1);
                        sm.q[slot].phase = ph; sm.q[slot].mt = mt; sm.q[slot].nt = nt; sm.q[slot].it = it;
                        mbar_arrive(&sm.q_full[slot]);
                        ++qn;
                    }
                    for (int seg = 0; seg < nseg; ++seg)
                        for (int kc = 0; kc < KC; ++kc) {
                            const uint32_t s = k % T_STAGES;
                            mbar_wait_wd(&sm.empty[s], ((k / T_STAGES) & 1) ^ 1);
                            unsigned char *st = stages + s * T_STAGE_BYTES;
                            mbar_expect_tx(&sm.full[s], T_STAGE_BYTES);
                            tma_load_2d(st + 0 * T_TILE_BYTES, a_hi[seg], &sm.full[s], kc * BK, a_row[seg] + mt * T_BM);
                            tma_load_2d(st + 1 * T_TILE_BYTES, a_lo[seg], &sm.full[s], kc * BK, a_row[seg] + mt * T_BM);
                            tma_load_2d(st + 2 * T_TILE_BYTES, w_hi, &sm.full[s], seg * kH + kc * BK, nt * T_BN);
                            tma_load_2d(st + 3 * T_TILE_BYTES, w_lo, &sm.full[s], seg * kH + kc * BK, nt * T_BN);
                            ++k;
                        }
                }
                if (!any) break;
            }
            const uint32_t slot = qn % D_Q;  // exit descriptor
            mbar_wait_wd(&sm.q_empty[slot], ((qn / D_Q) & 1) ^ 1);
            sm.q[slot].phase = -1;
            mbar_arrive(&sm.q_full[slot]);
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc_bf16(T_BM, T_BN);
            uint32_t k = 0, qn = 0, tile = 0;
            for (;;) {
                const uint32_t slot = qn % D_Q;
                mbar_wait_wd(&sm.q_full[slot], (qn / D_Q) & 1);
                const int ph = sm.q[slot].phase;
                mbar_arrive(&sm.q_empty[slot]);
                ++qn;
                if (ph < 0) break;
                const int nk = ph == PH_B ? 2 * KC : KC;
                const uint32_t buf = tile & 1, use = tile >> 1;
                mbar_wait_wd(&sm.acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * T_BN;
                for (int kc = 0; kc < nk; ++kc) {
                    const uint32_t s = k % T_STAGES;
                    mbar_wait_wd(&sm.full[s], (k / T_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t st = smem_u32(stages + s * T_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                        const uint32_t off = kk * UMMA_K * 2;
                        const uint64_t ah = make_sdesc_sw128(st + 0 * T_TILE_BYTES + off), al = make_sdesc_sw128(st + 1 * T_TILE_BYTES + off);
                        const uint64_t wh = make_sdesc_sw128(st + 2 * T_TILE_BYTES + off), wl = make_sdesc_sw128(st + 3 * T_TILE_BYTES + off);
                        umma_bf16(acc, al, wh, idesc, (kc | kk) != 0);
                        umma_bf16(acc, ah, wl, idesc, 1);
                        umma_bf16(acc, ah, wh, idesc, 1);
                    }
                    umma_commit(&sm.empty[s]);
                    ++k;
                }
                umma_commit(&sm.acc_full[buf]);
                ++tile;
            }
        }
    } else if (warp >= 4) {  // ===================== epilogue: 16 warps =====================
        const int e = warp - 4, q = e & 3, cgp = e >> 2;  // TMEM lane quarter (== warp % 4), column group
        const int etid = tid - 128;
        const int r_in = q * 32 + lane;
        uint32_t qn = 0, tile = 0;
        for (;;) {
            const uint32_t slot = qn % D_Q;
            mbar_wait_wd(&sm.q_full[slot], (qn / D_Q) & 1);
            const DfDesc d = sm.q[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.q_empty[slot]);
            ++qn;
            if (d.phase < 0) break;
            const int ph = d.phase, mt = d.mt, nt = d.nt, it = d.it, par = it & 1;
            const uint32_t buf = tile & 1, use = tile >> 1;
            const int row = mt * T_BM + r_in;
            const int nb = nt * T_BN + cgp * 32;  // first of this thread's 32 columns
            const TCtl c = load_ctl(p.ctl + row);
            mbar_wait_wd(&sm.acc_full[buf], use & 1);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld32(tmem_base + buf * T_BN + ((uint32_t)(q * 32) << 16) + cgp * 32, r);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sm.acc_empty[buf]);  // accumulator is in registers: the MMA warp may reuse the buffer
            ++tile;
            float best_v = -INFINITY;
            int best_i = 0x7fffffff;
            if (c.active) {
                if (ph == PH_A || ph == PH_B) {
                    const float *addp = ph == PH_A ? p.g0p + (size_t)c.last * kG + nb : p.b1p + nb;
                    float *cst = (ph == PH_A ? p.c0 : p.c1) + (size_t)row * kH + nb / 4;
                    float *hf = (ph == PH_A ? p.h0f : p.h1f) + (size_t)row * kH + nb / 4;
                    const size_t ob = ((size_t)(par ^ 1) * p.Mpad + row) * kH + nb / 4;
                    __nv_bfloat16 *bh = (ph == PH_A ? p.h0b_hi : p.h1b_hi) + ob, *bl = (ph == PH_A ? p.h0b_lo : p.h1b_lo) + ob;
                    float cold[8], hnew[8];
                    *reinterpret_cast<float4 *>(&cold[0]) = __ldcg(reinterpret_cast<const float4 *>(cst));
                    *reinterpret_cast<float4 *>(&cold[4]) = __ldcg(reinterpret_cast<const float4 *>(cst) + 1);
                    __align__(16) __nv_bfloat16 vh[8], vl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 ad = __ldg(reinterpret_cast<const float4 *>(addp) + j);
                        const float gi = sigm(__uint_as_float(r[4 * j + 0]) + ad.x), gf = sigm(__uint_as_float(r[4 * j + 1]) + ad.y);
                        const float gg = tanhf(__uint_as_float(r[4 * j + 2]) + ad.z), go = sigm(__uint_as_float(r[4 * j + 3]) + ad.w);
                        const float cn = gf * cold[j] + gi * gg;
                        cold[j] = cn;
                        hnew[j] = go * tanhf(cn);
                        split_bf16(hnew[j], vh[j], vl[j]);
                    }
                    reinterpret_cast<float4 *>(cst)[0] = *reinterpret_cast<float4 *>(&cold[0]);
                    reinterpret_cast<float4 *>(cst)[1] = *reinterpret_cast<float4 *>(&cold[4]);
                    reinterpret_cast<float4 *>(hf)[0] = *reinterpret_cast<float4 *>(&hnew[0]);
                    reinterpret_cast<float4 *>(hf)[1] = *reinterpret_cast<float4 *>(&hnew[4]);
                    *reinterpret_cast<uint4 *>(bh) = *reinterpret_cast<uint4 *>(vh);
                    *reinterpret_cast<uint4 *>(bl) = *reinterpret_cast<uint4 *>(vl);
                } else if (ph == PH_C) {
                    const float *ep = p.E + ((size_t)p.eoff[p.perm[row]] + c.t) * kH + nb;
                    __nv_bfloat16 *bh = p.zb_hi + (size_t)row * kH + nb, *bl = p.zb_lo + (size_t)row * kH + nb;
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        __align__(16) __nv_bfloat16 vh[8], vl[8];
                        const float4 e0 = __ldg(reinterpret_cast<const float4 *>(ep) + 2 * j8), e1 = __ldg(reinterpret_cast<const float4 *>(ep) + 2 * j8 + 1);
                        const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v = __uint_as_float(r[8 * j8 + j]) + ev[j];
                            split_bf16(p.relu ? fmaxf(v, 0.f) : tanhf(v), vh[j], vl[j]);
                        }
                        reinterpret_cast<uint4 *>(bh)[j8] = *reinterpret_cast<uint4 *>(vh);
                        reinterpret_cast<uint4 *>(bl)[j8] = *reinterpret_cast<uint4 *>(vl);
                    }
                } else {  // PH_D: first-max argmax over this thread's 32 columns (zero_copy.rs:190-232 tie rule)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nb + j;
                        if (n < kV) {
                            const float v = __uint_as_float(r[j]) + __ldg(p.boutp + n);
                            if (v > best_v || best_i == 0x7fffffff) { best_v = v; best_i = n; }
                        }
                    }
                    p.pval[(size_t)row * NPART + nt * 4 + cgp] = best_v;
                    p.pidx[(size_t)row * NPART + nt * 4 + cgp] = best_i;
                }
            }
            __threadfence();
            fence_proxy_async();  // these generic-proxy stores are read by other CTAs' TMA (async proxy)
            named_bar_sync(1, D_EPI_THREADS);
            if (ph != PH_D) {
                if (etid == 0) {
                    __threadfence();  // cumulative release of every epilogue thread's stores
                    atomicAdd(ph == PH_A ? p.cnt_a + mt : (ph == PH_B ? p.cnt_b + mt : p.cnt_c + mt), 1);
                }
                continue;
            }
            if (etid == 0) {
                const int old = atomicAdd(&p.done_cnt[mt], 1);
                sm.last_flag = (old == NT_O - 1);
                sm.act_cnt = 0;
            }
            named_bar_sync(1, D_EPI_THREADS);
            if (!sm.last_flag) continue;
            // ---- this CTA finished the M-tile's last vocab tile: per-stream control update (one thread per row) ----
            __threadfence();
            if (cgp == 0) {
                TCtl n = c;
                if (n.active) {
                    float bv = __ldcg(p.pval + (size_t)row * NPART);
                    int bi = __ldcg(p.pidx + (size_t)row * NPART);
                    for (int qi = 1; qi < NPART; ++qi) {
                        const int oi = __ldcg(p.pidx + (size_t)row * NPART + qi);
                        const float v = __ldcg(p.pval + (size_t)row * NPART + qi);
                        if (oi != 0x7fffffff && (bi == 0x7fffffff || v > bv)) { bv = v; bi = oi; }
                    }
                    const int len = p.lens[p.perm[row]];
                    n.nsteps += 1;                       // state carried unconditionally (decoder_optimized.rs:154)
                    n.sym += 1;                          // :133
                    if (bi == p.blank) {                 // :171-173
                        n.t += 1; n.sym = 0;
                        if (n.t >= len) n.active = 0;
                    } else {
                        p.tokens[(size_t)p.perm[row] * p.max_total + n.total] = bi;   // :176
                        n.total += 1;
                        n.last = bi;
                        if (n.total >= p.max_total) n.active = 0;                    // :179-188
                        else if (n.sym >= p.max_sym) {                               // :133-137
                            n.t += 1; n.sym = 0;
                            if (n.t >= len) n.active = 0;
                        }
                        if (n.active && bi >= kEmbRows) { n.active = 0; n.failed = 1; }  // next step would fail (:148-152)
                    }
                    p.ctl[row] = n;
                    if (n.active) atomicAdd(&sm.act_cnt, 1);
                }
            }
            __threadfence();
            named_bar_sync(1, D_EPI_THREADS);
            const int alive = sm.act_cnt;
            if (alive == 0) {  // M-tile finished: write its streams' results (all 512 threads)
                for (int rr = etid; rr < T_BM; rr += D_EPI_THREADS) {
                    const int grow = mt * T_BM + rr;
                    if (grow < p.B) {
                        const TCtl f = load_ctl(p.ctl + grow);
                        const int b = p.perm[grow];
                        p.ntok[b] = f.failed ? -1 : f.total;
                        if (p.nsteps) p.nsteps[b] = f.nsteps;
                        if (f.failed) atomicAdd(&p.total_active[2], 1);
                    }
                }
                if (p.s1 && p.s2) {
                    for (int i = etid; i < T_BM * kH; i += D_EPI_THREADS) {
                        const int grow = mt * T_BM + i / kH, j = i % kH;
                        if (grow < p.B) {
                            const int b = p.perm[grow];
                            const size_t src = (size_t)grow * kH + j;
                            p.s1[tc_state_off(p, 0, b) + j] = __ldcg(p.h0f + src);
                            p.s1[tc_state_off(p, 1, b) + j] = __ldcg(p.h1f + src);
                            p.s2[tc_state_off(p, 0, b) + j] = __ldcg(p.c0 + src);
                            p.s2[tc_state_off(p, 1, b) + j] = __ldcg(p.c1 + src);
                        }
                    }
                }
                __threadfence();
                named_bar_sync(1, D_EPI_THREADS);
            }
            if (etid == 0) {
                p.done_cnt[mt] = 0;
                if (alive == 0) p.dead_at[mt] = it + 1;
                __threadfence();
                atomicAdd(p.ctl_done + mt, 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(sm.tmem_slot, 2 * T_BN);
}

}  // namespace

// ---- host side of engine 2 ----
void decoder_tc_release(Ctx *c) {
    if (!c->dec || !c->dec->tc) return;
    TcWeights *w = c->dec->tc;
    for (__nv_bfloat16 *p : {w->whh0_hi, w->whh0_lo, w->w1_hi, w->w1_lo, w->wp_hi, w->wp_lo, w->wo_hi, w->wo_lo, w->we_hi, w->we_lo})
        if (p) cudaFree(p);
    delete w;
    c->dec->tc = nullptr;
}

cudaError_t decoder_tc_prepare_weights(Ctx *c) {
    DecoderPriv *d = c->dec;
    if (!d->tc) d->tc = new TcWeights();
    TcWeights *w = d->tc;
    const BlobLayout L = blob_layout();
    cudaError_t e;
    auto alloc = [&](__nv_bfloat16 **p, size_t n) -> cudaError_t { return *p ? cudaSuccess : cudaMalloc(p, sizeof(__nv_bfloat16) * n); };
    if ((e = alloc(&w->whh0_hi, (size_t)kG * kH)) != cudaSuccess || (e = alloc(&w->whh0_lo, (size_t)kG * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->w1_hi, (size_t)kG * 2 * kH)) != cudaSuccess || (e = alloc(&w->w1_lo, (size_t)kG * 2 * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->wp_hi, (size_t)kH * kH)) != cudaSuccess || (e = alloc(&w->wp_lo, (size_t)kH * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->wo_hi, (size_t)V_PAD_TC * kH)) != cudaSuccess || (e = alloc(&w->wo_lo, (size_t)V_PAD_TC * kH)) != cudaSuccess) return e;
    if ((e = alloc(&w->we_hi, (size_t)kH * kEnc)) != cudaSuccess || (e = alloc(&w->we_lo, (size_t)kH * kEnc)) != cudaSuccess) return e;
    cudaMemsetAsync(w->wo_hi, 0, sizeof(__nv_bfloat16) * (size_t)V_PAD_TC * kH, c->stream);
    cudaMemsetAsync(w->wo_lo, 0, sizeof(__nv_bfloat16) * (size_t)V_PAD_TC * kH, c->stream);
    if ((e = launch_split_rows(c, d->whh0p, kH, w->whh0_hi, w->whh0_lo, kH, kG, kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, d->w1p, 2 * kH, w->w1_hi, w->w1_lo, 2 * kH, kG, 2 * kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, c->w_blob + L.w_pred, kH, w->wp_hi, w->wp_lo, kH, kH, kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, c->w_blob + L.w_out, kH, w->wo_hi, w->wo_lo, kH, kV, kH)) != cudaSuccess) return e;
    if ((e = launch_split_rows(c, c->w_blob + L.w_enc, kEnc, w->we_hi, w->we_lo, kEnc, kH, kEnc)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_whh0_hi, w->whh0_hi, kG, kH, kH, T_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_whh0_lo, w->whh0_lo, kG, kH, kH, T_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_w1_hi, w->w1_hi, kG, 2 * kH, 2 * kH, T_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_w1_lo, w->w1_lo, kG, 2 * kH, 2 * kH, T_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_wp_hi, w->wp_hi, kH, kH, kH, T_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_wp_lo, w->wp_lo, kH, kH, kH, T_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_wo_hi, w->wo_hi, V_PAD_TC, kH, kH, T_BN)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->m_wo_lo, w->wo_lo, V_PAD_TC, kH, kH, T_BN)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(greedy_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(greedy_df_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM)) != cudaSuccess) return e;
    int nb = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, greedy_tc_kernel, T_THREADS, T_SMEM)) != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    w->coop_blocks_per_sm = 1;  // one CTA per SM: the smem ring and 256 TMEM columns are sized for that
    return decoder_ws_prepare(c);
}

static size_t tc_align(size_t x) { return (x + 1023) & ~(size_t)1023; }

cudaError_t launch_greedy_decode_tc(Ctx *c, const float *enc_dev, const float *enc_host, int B, int T, const int32_t *lens_dev,
                                    const int32_t *lens_host, const int32_t *slots_dev, float *s1_dev, float *s2_dev,
                                    int32_t *tokens_dev, int32_t *ntok_dev, int32_t *nsteps_dev, const int64_t *enc_off_host) {
    // enc_off_host != nullptr: packed encoder outputs — stream b is a [1024][lens[b]] block at enc + enc_off_host[b]
    DecoderPriv *d = c->dec;
    TcWeights *w = d->tc;
    const int Tq = T > 0 ? T : 1;
    const int MT = (B + T_BM - 1) / T_BM, Mpad = MT * T_BM;
    if (MT > MAX_MT) return cudaErrorInvalidValue;
    // decode_engine 0 (auto) / 4: weight-stationary dataflow kernel (decoder_ws.cu) when the device has the 147 SMs it needs
    const bool use_ws = (c->cfg.decode_engine == 0 || c->cfg.decode_engine == 4) && decoder_ws_supported(c);
    const size_t MH = (size_t)Mpad * kH;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += tc_align(bytes); return o; };
    const size_t oE = take(sizeof(float) * (size_t)B * Tq * kH);
    const size_t oeh = take(2 * (size_t)B * Tq * kEnc), oel = take(2 * (size_t)B * Tq * kEnc);
    const size_t meta_bytes = sizeof(long long) * ((size_t)B + 1) + sizeof(int) * ((size_t)Mpad + (size_t)B + 1);
    const size_t ometa = take(meta_bytes);  // src_off[B+1] (packed input), perm[Mpad], eoff[B+1]
    const size_t operm = ometa + sizeof(long long) * ((size_t)B + 1);
    size_t ows = 0, ws_bytes = 0;
    size_t oh0h = 0, oh0l = 0, oh1h = 0, oh1l = 0, ozh = 0, ozl = 0, oh0f = 0, oh1f = 0, oc0 = 0, oc1 = 0, opv = 0, opi = 0, octl = 0, ocnt = 0;
    cudaError_t e;
    if (use_ws) {
        if ((e = launch_greedy_ws(c, nullptr, B, T, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &ws_bytes)) != cudaSuccess) return e;
        ows = take(ws_bytes);
    } else {
        oh0h = take(2 * 2 * MH); oh0l = take(2 * 2 * MH); oh1h = take(2 * 2 * MH); oh1l = take(2 * 2 * MH);
        ozh = take(2 * MH); ozl = take(2 * MH);
        oh0f = take(4 * MH); oh1f = take(4 * MH); oc0 = take(4 * MH); oc1 = take(4 * MH);
        opv = take(sizeof(float) * (size_t)Mpad * NPART); opi = take(sizeof(int) * (size_t)Mpad * NPART);
        octl = take(sizeof(TCtl) * (size_t)Mpad);
        ocnt = take(sizeof(int) * (7 * (size_t)MT + 4));
    }
    if ((e = d->work.reserve(off)) != cudaSuccess) return e;
    char *base = d->work.as<char>();

    // rows sorted by encoded length (descending, stable): active rows stay a prefix, whole M-tiles retire early
    if ((e = c->pin[1].reserve(meta_bytes)) != cudaSuccess) return e;
    long long *h_soff = c->pin[1].as<long long>();
    for (int i = 0; i <= B; ++i) h_soff[i] = enc_off_host ? (long long)enc_off_host[i] : 0;
    int *h_perm = reinterpret_cast<int *>(h_soff + B + 1);
    int *h_eoff = h_perm + Mpad;  // first packed row of each stream's valid frames in E
    h_eoff[0] = 0;
    for (int i = 0; i < B; ++i) h_eoff[i + 1] = h_eoff[i] + lens_host[i];
    for (int i = 0; i < B; ++i) h_perm[i] = i;
    std::stable_sort(h_perm, h_perm + B, [&](int a, int b2) { return lens_host[a] > lens_host[b2]; });
    for (int i = B; i < Mpad; ++i) h_perm[i] = 0;
    if ((e = cudaMemcpyAsync(base + ometa, h_soff, meta_bytes, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
    const long long *soff_dev = enc_off_host ? reinterpret_cast<const long long *>(base + ometa) : nullptr;
    const int *eoff_dev = reinterpret_cast<int *>(base + operm) + Mpad;

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
    if (use_ws)
        return launch_greedy_ws(c, E, B, T, lens_dev, reinterpret_cast<int *>(base + operm), eoff_dev, slots_dev, s1_dev, s2_dev,
                                tokens_dev, ntok_dev, nsteps_dev, base + ows, &ws_bytes);

    if ((e = cudaMemsetAsync(base + ocnt, 0, sizeof(int) * (7 * (size_t)MT + 4), c->stream)) != cudaSuccess) return e;
    // activation buffers start defined (padding rows feed the MMA too)
    if ((e = cudaMemsetAsync(base + oh0h, 0, (ozl + tc_align(2 * MH)) - oh0h, c->stream)) != cudaSuccess) return e;

    TcDecParams p;
    std::memset(&p, 0, sizeof(p));
    p.h0b_hi = reinterpret_cast<__nv_bfloat16 *>(base + oh0h); p.h0b_lo = reinterpret_cast<__nv_bfloat16 *>(base + oh0l);
    p.h1b_hi = reinterpret_cast<__nv_bfloat16 *>(base + oh1h); p.h1b_lo = reinterpret_cast<__nv_bfloat16 *>(base + oh1l);
    p.zb_hi = reinterpret_cast<__nv_bfloat16 *>(base + ozh); p.zb_lo = reinterpret_cast<__nv_bfloat16 *>(base + ozl);
    if ((e = make_tmap_bf16(&p.h0_hi, p.h0b_hi, 2 * (uint64_t)Mpad, kH, kH, T_BM)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.h0_lo, p.h0b_lo, 2 * (uint64_t)Mpad, kH, kH, T_BM)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.h1_hi, p.h1b_hi, 2 * (uint64_t)Mpad, kH, kH, T_BM)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.h1_lo, p.h1b_lo, 2 * (uint64_t)Mpad, kH, kH, T_BM)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.z_hi, p.zb_hi, (uint64_t)Mpad, kH, kH, T_BM)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.z_lo, p.zb_lo, (uint64_t)Mpad, kH, kH, T_BM)) != cudaSuccess) return e;
    p.whh0_hi = w->m_whh0_hi; p.whh0_lo = w->m_whh0_lo; p.w1_hi = w->m_w1_hi; p.w1_lo = w->m_w1_lo;
    p.wp_hi = w->m_wp_hi; p.wp_lo = w->m_wp_lo; p.wo_hi = w->m_wo_hi; p.wo_lo = w->m_wo_lo;
    p.g0p = d->g0p; p.b1p = d->b1p; p.boutp = d->boutp; p.E = E;
    p.B = B; p.Mpad = Mpad; p.MT = MT; p.T = Tq;
    p.lens = lens_dev; p.slots = slots_dev; p.perm = reinterpret_cast<int *>(base + operm); p.eoff = eoff_dev;
    p.h0f = reinterpret_cast<float *>(base + oh0f); p.h1f = reinterpret_cast<float *>(base + oh1f);
    p.c0 = reinterpret_cast<float *>(base + oc0); p.c1 = reinterpret_cast<float *>(base + oc1);
    p.pval = reinterpret_cast<float *>(base + opv); p.pidx = reinterpret_cast<int *>(base + opi);
    p.ctl = reinterpret_cast<TCtl *>(base + octl);
    int *cnt = reinterpret_cast<int *>(base + ocnt);
    p.tile_active = cnt; p.done_cnt = cnt + MT; p.total_active = cnt + 2 * MT;
    p.cnt_a = cnt + 2 * MT + 4; p.cnt_b = p.cnt_a + MT; p.cnt_c = p.cnt_b + MT; p.ctl_done = p.cnt_c + MT; p.dead_at = p.ctl_done + MT;
    if (slots_dev) { p.s1 = c->slot_s1; p.s2 = c->slot_s2; } else { p.s1 = s1_dev; p.s2 = s2_dev; }
    p.tokens = tokens_dev; p.ntok = ntok_dev; p.nsteps = nsteps_dev;
    p.max_sym = c->cfg.max_symbols_per_step; p.max_total = c->cfg.max_total_tokens; p.blank = c->cfg.blank_id;
    p.relu = c->cfg.joint_activation;
    d->fail_count_dev = p.total_active + 2;

    void *params[] = {&p};
    ProfScope prof(c, PK_GREEDY);
    if (c->cfg.decode_engine == 2) {  // grid-synchronised variant
        int grid = std::min(w->coop_blocks_per_sm * c->sm_count, MT * NT_G);
        if (grid < 1) grid = 1;
        e = cudaLaunchCooperativeKernel((const void *)greedy_tc_kernel, dim3(grid), dim3(T_THREADS), params, T_SMEM, c->stream);
    } else {                          // dataflow variant (default): one CTA per SM, all co-resident
        if (MT > 255) return cudaErrorInvalidValue;
        const int grid = std::max(1, std::min(c->sm_count, MT * D_TILES_PER_MT));
        e = cudaLaunchCooperativeKernel((const void *)greedy_df_kernel, dim3(grid), dim3(D_THREADS), params, D_SMEM, c->stream);
    }
    c->launches++;
    return e;
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
