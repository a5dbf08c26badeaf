// bench_pipe.cu — microbenchmark (not part of the library): the decode unit's inner pipeline in isolation.  A TMA thread streams
// [128 rows][64 k] bf16 tiles of an L2-resident tensor through a ring of NRING x 16 KB; an MMA thread consumes every slot with
// four tcgen05.mma (M = 128, N = 128, K = 16) whose A operand sits in tensor memory (mode 1) or in shared memory (mode 0), then
// commits the slot free.  Prints the time per "unit" (20 slots = one 128 x 640 hi/lo activation tile) with all 148 SMs running.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_pipe bench_pipe.cu -lcuda && ./bench_pipe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int UNIT = 128 * 64 * 2, KC = 10, MAXR = 12;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) { while (!mbar_try(b, par)) {} }
__device__ __forceinline__ void tma_load(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) { return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }

__global__ void __launch_bounds__(128, 1) pipe_kernel(const __grid_constant__ CUtensorMap map, int mode, int nring, int n_units, int n_tiles, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[MAXR], empty[MAXR], done;
    __shared__ uint32_t tslot;
    if (threadIdx.x == 0) {
        for (int s = 0; s < MAXR; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    unsigned char *ring = smem + (mode == 0 ? 10 * UNIT : 0);  // mode 0: a resident 160 KB "weight" region in front, as in the library
    const long long t0 = clock64();
    const int slots = n_units * 2 * KC;
    if (threadIdx.x == 0 && mode != 2) {
        for (int u = 0; u < slots; ++u) {
            const int s = u % nring, tile = (u / (2 * KC)) % n_tiles, kc = (u >> 1) % KC;
            mbar_wait(&empty[s], ((u / nring) & 1) ^ 1);
            mbar_expect_tx(&full[s], UNIT);
            tma_load(ring + s * UNIT, &map, &full[s], kc * 64, ((u & 1) * n_tiles + tile) * 128);
        }
    } else if (threadIdx.x == 32) {
        const uint32_t id128 = idesc_bf16(128, 128), id64 = idesc_bf16(128, 64);
        for (int u = 0; u < slots; ++u) {
            const int s = u % nring, kc = (u >> 1) % KC;
            if (mode != 2) mbar_wait(&full[s], (u / nring) & 1);
            const uint32_t slot = smem_u32(ring + s * UNIT);
            if (mode >= 1) {  // A (weights) in TMEM columns [0, 320), B = slot, N = 128
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm + 320),
                                 "r"(tm + kc * 32 + kk * 8), "l"(sdesc(slot + kk * 32)), "r"(id128), "r"(1u) : "memory");
            } else {          // library's shared-memory-stationary pattern: A = slot, B = resident chunk; N = 128 for hi slots, 64 for lo
                const uint32_t w = smem_u32(smem + kc * UNIT);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
                                 "l"(sdesc(slot + kk * 32)), "l"(sdesc(w + kk * 32)), "r"((u & 1) ? id64 : id128), "r"(1u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
        mbar_wait(&done, 0);
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    const int n_tiles = 8, rows = 2 * n_tiles * 128, cols = 640;
    __nv_bfloat16 *d;
    CK(cudaMalloc(&d, sizeof(__nv_bfloat16) * rows * cols));
    CK(cudaMemset(d, 0, sizeof(__nv_bfloat16) * rows * cols));
    long long *cyc, h[148];
    CK(cudaMalloc(&cyc, sizeof(long long) * 148));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    if (reinterpret_cast<EncodeFn>(fp)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    const int smem = 14 * UNIT;
    CK(cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int n_units = 400;
    struct { int mode, nring; const char *what; } cases[] = {{0, 4, "weights in smem, ring 4 (library, shared-memory-stationary)"},
        {1, 4, "weights in TMEM, ring 4"}, {1, 9, "weights in TMEM, ring 9"}, {2, 1, "weights in TMEM, NO TMA, same slot"}, {2, 9, "weights in TMEM, NO TMA, 9 slots"}};
    for (auto &c : cases)
        for (int grid : {1, 40, 148}) {
            pipe_kernel<<<grid, 128, smem>>>(map, c.mode, c.nring, n_units, n_tiles, cyc);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            double mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? (double)h[i] : mx;
            printf("%-62s grid %3d: %.0f cycles = %.2f us per unit (327 KB ingest, 80 mma)\n", c.what, grid, mx / n_units, mx / n_units / 1965.0);
        }
    return 0;
}
