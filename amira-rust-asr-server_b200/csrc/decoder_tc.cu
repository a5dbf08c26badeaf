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

// encoder_outputs [B][1024][T] (t contiguous) -> K-major rows [(b*T + t)][1024] hi/lo bf16, via a 32x32 smem transpose
__global__ void split_transpose_enc_kernel(const float *__restrict__ enc, int T, const int *__restrict__ lens,
                                           __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    if (t0 >= lens[b]) return;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    const float *src = enc + (size_t)b * kEnc * T;
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + tx;
        tile[j][tx] = t < T ? src[(size_t)(f0 + j) * T + t] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + j;
        if (t < T) {
            __nv_bfloat16 h, l;
            split_bf16(tile[tx][j], h, l);
            const size_t o = ((size_t)b * T + t) * kEnc + f0 + tx;
            hi[o] = h;
            lo[o] = l;
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

__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + G_STAGES * G_STAGE_BYTES);
    uint64_t *empty = full + G_STAGES;
    uint64_t *acc_full = empty + G_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * G_BM, n0 = blockIdx.x * G_BN;
    const int nk = (p.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&p.a_hi);
        tma_prefetch_desc(&p.a_lo);
        tma_prefetch_desc(&p.w_hi);
        tma_prefetch_desc(&p.w_lo);
    }
    if (warp == 1) tmem_alloc(tmem_slot, G_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % G_STAGES;
                mbar_wait(&empty[s], ((kc / G_STAGES) & 1) ^ 1);
                unsigned char *st = smem + s * G_STAGE_BYTES;
                mbar_expect_tx(&full[s], G_STAGE_BYTES);
                tma_load_2d(st + 0 * G_TILE_BYTES, &p.a_hi, &full[s], kc * BK, m0);
                tma_load_2d(st + 1 * G_TILE_BYTES, &p.a_lo, &full[s], kc * BK, m0);
                tma_load_2d(st + 2 * G_TILE_BYTES, &p.w_hi, &full[s], kc * BK, n0);
                tma_load_2d(st + 3 * G_TILE_BYTES, &p.w_lo, &full[s], kc * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_bf16(G_BM, G_BN);
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % G_STAGES;
                mbar_wait(&full[s], (kc / G_STAGES) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * G_STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                    const uint32_t off = kk * UMMA_K * 2;  // bytes along the 128-byte swizzle row
                    const uint64_t ah = make_sdesc_sw128(st + 0 * G_TILE_BYTES + off), al = make_sdesc_sw128(st + 1 * G_TILE_BYTES + off);
                    const uint64_t wh = make_sdesc_sw128(st + 2 * G_TILE_BYTES + off), wl = make_sdesc_sw128(st + 3 * G_TILE_BYTES + off);
                    umma_bf16(tmem_acc, al, wh, idesc, (kc | kk) != 0);
                    umma_bf16(tmem_acc, ah, wl, idesc, 1);
                    umma_bf16(tmem_acc, ah, wh, idesc, 1);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(acc_full);
        }
    } else {  // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < G_BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + c * 32, r);
            tmem_ld_wait();
            if (row < p.M) {
                float *dst = p.C + (size_t)row * p.ldc + n0 + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c * 32 + j;
                    if (n < p.N) dst[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n] : 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, G_BN);
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
    static bool attr_done = false;
    if (!attr_done) {
        if ((e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM)) != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((N + G_BN - 1) / G_BN, (M + G_BM - 1) / G_BM);
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

cudaError_t launch_split_transpose_enc(Ctx *c, const float *enc, int B, int T, const int *lens_dev, __nv_bfloat16 *hi,
                                       __nv_bfloat16 *lo) {
    if (B <= 0 || T <= 0) return cudaSuccess;
    dim3 grid((T + 31) / 32, kEnc / 32, B), block(32, 8);
    split_transpose_enc_kernel<<<grid, block, 0, c->stream>>>(enc, T, lens_dev, hi, lo);
    c->launches++;
    return cudaGetLastError();
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
