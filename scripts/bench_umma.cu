// bench_umma.cu — microbenchmark (not part of the library): cycles per tcgen05.mma (kind::f16, bf16, SS mode, K=16) issued
// back to back by one thread from resident shared-memory operands, for M=128 and N in {64, 128, 256}; and the same
// with the split-bf16 issue pattern of decoder_ws.cu (three products per k-step re-reading the A tile).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_umma bench_umma.cu && ./bench_umma
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) { return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }

// mode 0: n_mma identical UMMAs of N columns.  mode 1: decoder_ws pattern (per k-step: a_hi*w_hi, a_hi*w_lo, a_lo*w_hi, N = 64).
// mode 2: fused pattern (per k-step: a_hi*[w_hi;w_lo] with N = 128, a_lo*w_hi with N = 64).
__global__ void __launch_bounds__(128, 1) umma_kernel(int mode, int N, int n_iter, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[8];
    __shared__ uint32_t tslot;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    const uint32_t acc = tslot;
    if (threadIdx.x == 0) {
        const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + 16384, w = a_hi + 32768;  // w: up to 256 rows x 128 B = 32 KB
        const long long t0 = clock64();
        if (mode == 0) {
            const uint32_t id = idesc_bf16(128, N);
            for (int i = 0; i < n_iter; ++i)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_hi + kk * 32), sdesc(w + kk * 32), id, 1);
        } else if (mode == 1) {
            const uint32_t id = idesc_bf16(128, 64);
            for (int i = 0; i < n_iter; ++i) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_hi + kk * 32), sdesc(w + kk * 32), id, 1);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_hi + kk * 32), sdesc(w + 8192 + kk * 32), id, 1);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_lo + kk * 32), sdesc(w + kk * 32), id, 1);
            }
        } else if (mode == 2) {
            const uint32_t id128 = idesc_bf16(128, 128), id64 = idesc_bf16(128, 64);
            for (int i = 0; i < n_iter; ++i) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_hi + kk * 32), sdesc(w + kk * 32), id128, 1);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_lo + kk * 32), sdesc(w + kk * 32), id64, 1);
            }
        } else if (mode >= 5 && mode <= 7) {  // A in TMEM, N = 128; mode 5: B walks 9 different 16 KB tiles; 6: A walks 10 k-chunks; 7: both
            const uint32_t id = idesc_bf16(128, 128);
            for (int i = 0; i < n_iter; ++i) {
                const uint32_t bt = (mode == 5 || mode == 7) ? a_hi + (i % 9) * 16384 : w;
                const uint32_t at = acc + 128 + ((mode == 6 || mode == 7) ? (i % 10) * 32 : 0);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t bd = sdesc(bt + kk * 32);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(acc),
                                 "r"(at + kk * 8), "l"(bd), "r"(id), "r"(1u) : "memory");
                }
            }
        } else if (mode == 4 || mode == 8) {  // A operand in tensor memory (columns 256..), B from shared memory, N columns; mode 8: M = 64
            const uint32_t id = idesc_bf16(mode == 8 ? 64 : 128, N);
            for (int i = 0; i < n_iter; ++i)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t bd = sdesc(w + kk * 32);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(acc),
                                 "r"(acc + 256 + kk * 8), "l"(bd), "r"(id), "r"(1u) : "memory");
                }
        } else {  // mode 3: fused pattern with a tcgen05.commit after every group of four (as a smem ring would need)
            const uint32_t id128 = idesc_bf16(128, 128), id64 = idesc_bf16(128, 64);
            for (int i = 0; i < n_iter; ++i) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_hi + kk * 32), sdesc(w + kk * 32), id128, 1);
                commit(&bar2[(2 * i) & 7]);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma(acc, sdesc(a_lo + kk * 32), sdesc(w + kk * 32), id64, 1);
                commit(&bar2[(2 * i + 1) & 7]);
            }
        }
        commit(&bar);
        while (!mbar_try(&bar, 0)) {}
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(acc) : "memory");
}

int main() {
    long long *out, h[148];
    CK(cudaMalloc(&out, sizeof(long long) * 148));
    CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608));
    const int n_iter = 2000;
    struct { int mode, N; const char *what; int mmas; double units; } cases[] = {
        {0, 64, "N=64", 4, 0}, {0, 128, "N=128", 4, 0}, {0, 256, "N=256", 4, 0},
        {1, 64, "split pattern 3 x N=64 per k-step", 12, 0}, {2, 128, "fused pattern N=128 + N=64 per k-step", 8, 0},
        {3, 128, "fused pattern + commit per 4 mma", 8, 0}, {4, 64, "A in TMEM, N=64", 4, 0}, {4, 128, "A in TMEM, N=128", 4, 0},
        {4, 256, "A in TMEM, N=256", 4, 0}, {5, 128, "A in TMEM fixed, B walks 9 tiles", 4, 0}, {6, 128, "A in TMEM walks 10 chunks, B fixed", 4, 0},
        {7, 128, "A in TMEM walks, B walks", 4, 0}, {8, 128, "A in TMEM, M=64, N=128", 4, 0}, {8, 64, "A in TMEM, M=64, N=64", 4, 0}, {8, 32, "A in TMEM, M=64, N=32", 4, 0}, {4, 32, "A in TMEM, M=128, N=32", 4, 0}};
    for (auto &c : cases)
        for (int grid : {1, 148}) {
            umma_kernel<<<grid, 128, 196608>>>(c.mode, c.N, n_iter, out);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            double mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? (double)h[i] : mx;
            printf("%-42s grid %3d: %.1f cycles per tcgen05.mma, %.1f cycles per 64-wide k-chunk group (%d mma)\n", c.what, grid,
                   mx / (n_iter * (double)c.mmas), mx / n_iter, c.mmas);
        }
    return 0;
}
