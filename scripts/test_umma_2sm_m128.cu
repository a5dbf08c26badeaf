// test_umma_2sm_m128.cu — probe (not part of the library; derived from test_umma_2sm.cu): the SAME operands issued as M = 128
// (64 rows per CTA) into a separate 64-column accumulator: where do the rows / columns of D land, which TMEM lanes of A are read,
// and how many clocks does an instruction take?
// test_umma_2sm.cu — probe (not part of the library): tcgen05.mma.cta_group::2 with the A operand in TENSOR MEMORY.
// A CTA pair (cluster of 2): CTA r holds rows [128 r, 128 r + 128) of A[256][K] in its own TMEM and rows [64 r, 64 r + 64) of
// B[128][K] in its own shared memory (K-major, 128B swizzle, loaded by TMA with the completion signalled on the LEADER's mbarrier).
// The leader issues K/16 MMAs (M = 256, N = 128), commits with a multicast arrive to both CTAs; each CTA reads its half of D.
// Checks D = A B^T and which CTA's half of B lands in which columns, then times the issue rate of back-to-back MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o test_umma_2sm test_umma_2sm.cu -lcuda && ./test_umma_2sm
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int M = 256, N = 128, K = 640, KC = K / 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) { return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t a, uint32_t cta) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct Out { float *D; long long *clk; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmB, const __nv_bfloat16 *A, Out o, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];  // B half: KC chunks of [64 rows][64 k] bf16 = 8 KB each
    __shared__ uint64_t full_bar, done_bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full_bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tslot;
    const uint32_t a_tmem = tbase, d_tmem = tbase + 448;  // A: columns [0, 320); D (M = 128 form): [448, 512)
    {   // this CTA's 128 rows of A into its TMEM: lane = row, column j = (A[m][2j], A[m][2j+1])
        const int m = rank * 128 + warp * 32 + lane;
        for (int blk = 0; blk < K / 64; ++blk) {
            uint32_t r[32];
            for (int j = 0; j < 32; ++j) r[j] = *reinterpret_cast<const uint32_t *>(A + (size_t)m * K + blk * 64 + 2 * j);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(a_tmem + ((uint32_t)(warp * 32) << 16) + blk * 32),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
                "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // both TMEMs hold A before the leader issues
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // B: each CTA loads its 64 rows, chunk by chunk; completion on the LEADER's barrier
    if (tid == 0) {
        const uint32_t lead_bar = map_to_cta(smem_u32(&full_bar), 0);
        if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar)), "r"((uint32_t)(2 * KC * 8192)) : "memory");
        for (int kc = 0; kc < KC; ++kc)
            asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                             smem_u32(smem + kc * 8192)),
                         "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(lead_bar), "r"(kc * 64), "r"((int)rank * 64)
                         : "memory");
    }
    long long t0 = 0, t1 = 0;
    if (rank == 0 && tid == 0) {
        mbar_wait(&full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t id = idesc_bf16(128, N);
        t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            for (int kc = 0; kc < KC; ++kc)
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t bd = sdesc(smem_u32(smem + kc * 8192) + kk * 32);
                    const uint32_t acc = (rep | kc | kk) != 0;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
                                 "r"(a_tmem + kc * 32 + kk * 8), "l"(bd), "r"(id), "r"(acc)
                                 : "memory");
                }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(&done_bar, 0);
    if (rank == 0 && tid == 0) { t1 = clock64(); o.clk[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int m = rank * 128 + warp * 32 + lane;
        for (int c0 = 0; c0 < 64; c0 += 32) {
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
            for (int j = 0; j < 32; ++j) o.D[(size_t)m * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
    std::vector<float> fA((size_t)M * K), fB((size_t)N * K), ref((size_t)M * N), got((size_t)M * N);
    srand(7);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
            ref[(size_t)m * N + n] = (float)s;
        }
    __nv_bfloat16 *dA, *dB;
    Out o;
    CK(cudaMalloc(&dA, sizeof(__nv_bfloat16) * hA.size()));
    CK(cudaMalloc(&dB, sizeof(__nv_bfloat16) * hB.size()));
    CK(cudaMalloc(&o.D, sizeof(float) * got.size()));
    CK(cudaMalloc(&o.clk, sizeof(long long) * 4));
    CK(cudaMemcpy(dA, hA.data(), sizeof(__nv_bfloat16) * hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(__nv_bfloat16) * hB.size(), cudaMemcpyHostToDevice));
    CUtensorMap tm;
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N}, strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
    }
    const int smem_bytes = KC * 8192 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int rc = 0;
    for (int reps : {1, 8, 64}) {
        CK(cudaMemset(o.D, 0, sizeof(float) * got.size()));
        probe<<<2, 128, smem_bytes>>>(tm, dA, o, reps);
        CK(cudaDeviceSynchronize());
        long long clk = 0;
        CK(cudaMemcpy(&clk, o.clk, sizeof(clk), cudaMemcpyDeviceToHost));
        if (reps == 1) {
            CK(cudaMemcpy(got.data(), o.D, sizeof(float) * got.size(), cudaMemcpyDeviceToHost));
            // got[(cta*128 + lane) * N + col] for col < 64: find the unique (m, n) of ref with the same value
            for (int cta = 0; cta < 2; ++cta)
                for (int lane = 0; lane < 128; lane += (lane % 32 == 0 || lane % 32 == 30) ? 1 : 15) {
                    int ms[64], ns[64];
                    for (int col = 0; col < 64; ++col) {
                        const float v = got[(size_t)(cta * 128 + lane) * N + col];
                        int fm = -1, fn = -1, cnt = 0;
                        for (int m = 0; m < M; ++m)
                            for (int n = 0; n < N; ++n)
                                if (std::fabs((double)v - ref[(size_t)m * N + n]) < 1e-4 * (1.0 + std::fabs(v))) { fm = m; fn = n; ++cnt; }
                        ms[col] = cnt == 1 ? fm : (cnt == 0 ? -1 : -2);
                        ns[col] = cnt == 1 ? fn : (cnt == 0 ? -1 : -2);
                    }
                    printf("cta %d lane %3d: A rows (global) of cols 0,1,2,31,32,63: %d %d %d %d %d %d | B rows: %d %d %d %d %d %d\n", cta, lane, ms[0], ms[1], ms[2], ms[31], ms[32],
                           ms[63], ns[0], ns[1], ns[2], ns[31], ns[32], ns[63]);
                }
            double e_direct = 0;
            rc = e_direct < 2e-3 ? 0 : 1;
        }
        printf("  reps %d: %d MMAs (M=128 N=128 K=16, cta_group::2) issue -> commit %lld clk = %.1f clk per MMA\n", reps, reps * KC * 4, clk,
               (double)clk / (reps * KC * 4));
    }
    return rc;
}
