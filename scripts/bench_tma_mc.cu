// bench_tma_mc.cu — microbenchmark (not part of the library): per-SM TMA ingest rate from L2 with and without cluster
// multicast.  Every CTA streams the same L2-resident [rows][640] bf16 tensor through a 4 x 16 KB shared-memory ring
// (box {64 k, 128 rows}, 128B swizzle).  With cluster size C > 1 each CTA issues 1/C of every box (128/C rows) with
// .multicast::cluster so the box lands in all C CTAs.  Prints GB/s landed per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_tma_mc bench_tma_mc.cu -lcuda && ./bench_tma_mc
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int RING = 4, UNIT = 128 * 64 * 2, KC = 10;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) { while (!mbar_try(b, par)) {} }
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *b, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void tma_load(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_mc(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                     smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CSZ>
__global__ void __launch_bounds__(64, 1) ingest_kernel(const __grid_constant__ CUtensorMap map, int n_tiles, int n_rep, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + RING * UNIT), *empty = full + RING;
    const uint32_t rank = CSZ > 1 ? cluster_rank() : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RING; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CSZ); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CSZ > 1) cluster_sync();
    const long long t0 = clock64();
    const int units = n_rep * n_tiles * KC;
    if (threadIdx.x == 0) {  // producer
        for (int u = 0; u < units; ++u) {
            const int s = u % RING, tile = (u / KC) % n_tiles, kc = u % KC;
            mbar_wait(&empty[s], ((u / RING) & 1) ^ 1);
            mbar_expect_tx(&full[s], UNIT);
            if (CSZ == 1) tma_load(smem + s * UNIT, &map, &full[s], kc * 64, tile * 128);
            else tma_load_mc(smem + s * UNIT + rank * (UNIT / CSZ), &map, &full[s], kc * 64, tile * 128 + rank * (128 / CSZ), (uint16_t)((1u << CSZ) - 1));
        }
    } else if (threadIdx.x == 32) {  // consumer: frees the slot in every CTA of the cluster
        for (int u = 0; u < units; ++u) {
            const int s = u % RING;
            mbar_wait(&full[s], (u / RING) & 1);
            if (CSZ == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
            else for (uint32_t r = 0; r < CSZ; ++r) mbar_arrive_remote(&empty[s], r);
        }
    }
    __syncthreads();
    if (CSZ > 1) cluster_sync();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int CSZ>
void run(const CUtensorMap &map, int grid, int n_tiles, int n_rep, long long *cyc_dev) {
    const int smem = RING * UNIT + 256;
    CK(cudaFuncSetAttribute(ingest_kernel<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid / CSZ * CSZ);
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) {
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, ingest_kernel<CSZ>, map, n_tiles, n_rep, cyc_dev));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)n_rep * n_tiles * KC * UNIT;
    printf("cluster %d, grid %3d, %d tiles: %.3f ms -> %.1f GB/s landed per SM, %.2f TB/s landed chip-wide (L2 reads %.2f TB/s)\n", CSZ,
           (int)cfg.gridDim.x, n_tiles, ms, bytes / ms / 1e6, bytes * cfg.gridDim.x / ms / 1e9, bytes * cfg.gridDim.x / CSZ / ms / 1e9);
}

int main() {
    const int rows = 1024, cols = 640;
    __nv_bfloat16 *d;
    CK(cudaMalloc(&d, sizeof(__nv_bfloat16) * rows * cols));
    CK(cudaMemset(d, 0, sizeof(__nv_bfloat16) * rows * cols));
    long long *cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * 256));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
    for (int csz : {1, 2, 4, 8}) {
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        const cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)(128 / csz)};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        for (int grid : {8, 40, 144}) {
            if (csz == 1) run<1>(map, grid, 8, 50, cyc);
            if (csz == 2) run<2>(map, grid, 8, 50, cyc);
            if (csz == 4) run<4>(map, grid, 8, 50, cyc);
            if (csz == 8) run<8>(map, grid, 8, 50, cyc);
        }
    }
    return 0;
}
