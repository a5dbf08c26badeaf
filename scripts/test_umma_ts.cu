// test_umma_ts.cu — probe (not part of the library): tcgen05.mma with the A operand in TENSOR MEMORY (weights-in-TMEM idea).
// Writes A[128][K] (bf16) into TMEM with tcgen05.st under the hypothesis "row m -> lane m, k-pair j -> 32-bit column j",
// B[N][K] into shared memory in the K-major 128B-swizzled layout the library uses, runs K/16 MMAs and checks D = A B^T.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o test_umma_ts test_umma_ts.cu && ./test_umma_ts
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#ifndef MM
#define MM 128
#endif
constexpr int M = MM, N = 128, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) { return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D) {
    extern __shared__ __align__(1024) unsigned char smem[];  // B tile: [N rows][64 k] bf16, 128B swizzle
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B into the canonical layout: row n at n*128 B, 16-byte chunk c (8 k) stored at chunk (c ^ (n & 7))
    for (int i = tid; i < N * 8; i += 128) {
        const int n = i >> 3, c = i & 7;
        const uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)n * K + c * 8);
        *reinterpret_cast<uint4 *>(smem + n * 128 + ((c ^ (n & 7)) << 4)) = v;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tslot;
    const uint32_t a_tmem = tbase + 256, d_tmem = tbase;  // D: columns [0, N); A: columns [256, 256 + K/2)
    // A into TMEM: thread (warp w, lane l) owns TMEM lane 32 w + l = row m; column j holds (A[m][2j], A[m][2j+1])
    {
        const int m = warp * 32 + lane;
        uint32_t r[32];
        for (int j = 0; j < K / 2; ++j) r[j] = m < M ? *reinterpret_cast<const uint32_t *>(A + (size_t)m * K + 2 * j) : 0u;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(a_tmem + ((uint32_t)(warp * 32) << 16)),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
            "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
            "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // generic smem writes of B -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t id = idesc_bf16(M, N);
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint64_t bd = sdesc(smem_u32(smem) + kk * 32);
            const uint32_t acc = kk != 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
                         "r"(a_tmem + kk * 8), "l"(bd), "r"(id), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // wait, read D
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(d_tmem + ((uint32_t)(warp * 32) << 16) + c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) if (m < M) D[(size_t)m * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), got(M * N);
    srand(7);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
            ref[m * N + n] = (float)s;
        }
    __nv_bfloat16 *dA, *dB;
    float *dD;
    CK(cudaMalloc(&dA, sizeof(__nv_bfloat16) * M * K));
    CK(cudaMalloc(&dB, sizeof(__nv_bfloat16) * N * K));
    CK(cudaMalloc(&dD, sizeof(float) * M * N));
    CK(cudaMemcpy(dA, hA.data(), sizeof(__nv_bfloat16) * M * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(__nv_bfloat16) * N * K, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, sizeof(float) * M * N));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    probe<<<1, 128, 32768>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < M * N; ++i) maxerr = std::fmax(maxerr, std::fabs((double)got[i] - ref[i]));
    printf("A-in-TMEM MMA (M=%d N=%d K=%d): max abs err %.3e  (D[0][0]=%.4f ref %.4f, D[5][7]=%.4f ref %.4f, D[127][127]=%.4f ref %.4f)\n", M, N, K,
           maxerr, got[0], ref[0], got[5 * N + 7], ref[5 * N + 7], got[M * N - 1], ref[M * N - 1]);
    return maxerr < 1e-3 ? 0 : 1;
}
