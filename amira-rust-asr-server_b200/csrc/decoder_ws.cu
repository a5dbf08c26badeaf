// decoder_ws.cu — weight-stationary dataflow greedy decode on tcgen05 (decode_engine = 4, the default on B200).
//
// Same algebra, control flow and split-bf16 arithmetic as decoder_tc.cu (citations there and in decoder.cu: the loop is
// src/asr/decoder_optimized.rs:54-200, the step is src/asr/pipeline.rs:323-348 + src/triton/model.rs:581-722), but the
// work is laid out around one observation: the decoder weights split into exactly 147 slices of 64 output features x
// 640 inputs (layer-0 recurrent 40, layer-1 input 40, layer-1 recurrent 40, prediction projection 10, vocabulary 17),
// and one slice as split bf16 (hi + lo, 160 KB) fits the shared memory of one SM.  So:
//   * CTA s owns slice s for the whole kernel: its weights are loaded into shared memory ONCE (TMA, 128B swizzle) and
//     are never read from L2 again.  Only activations stream: per (128-stream M-tile, decode step) a CTA pulls the
//     128 x 640 hi/lo activation tile through a 4 x 16 KB TMA ring (measured: that ring sustains the SM's ~131 GB/s L2
//     ingest port) and issues 80 tcgen05.mma per unit: per k-step a_hi x [w_hi ; w_lo] (M=128, N=128: w_lo sits directly
//     below w_hi, so one instruction yields both products in two 64-column halves of the accumulator) and a_lo x w_hi
//     (N=64, into the first half); the epilogue adds the halves.  Four 128-column accumulators live in TMEM.
//   * Every CTA does exactly one unit of work per (M-tile, step): the schedule is static, perfectly balanced, and each
//     phase of a step runs on all of its slices' SMs at once (40 SMs per LSTM layer GEMM), which keeps the per-step
//     dependency chain short: layer-0 epilogue -> layer-1 input GEMM -> joint hidden GEMM -> vocabulary GEMM -> control.
//   * The two recurrent contractions (h0(t-1) W_hh0 and h1(t-1) W_hh1) do not depend on the token emitted at t-1, so
//     their CTAs run them one step ahead: layer-0 CTAs hold the accumulator in TMEM until the control update of the
//     previous step publishes the token (the epilogue adds the G0[token] row), layer-1 recurrent CTAs publish fp32
//     partial sums that the layer-1 input CTAs add in their epilogue.  Neither is on the critical path.
//   * Dependencies are per-M-tile monotonic counters in global memory (release: __threadfence + atomicAdd by the
//     epilogue; acquire: ld.acquire spin by the producer thread, then fence.proxy.async before the TMA loads), so
//     M-tiles advance independently; rows are sorted by encoded length so whole M-tiles retire early.
// Spin loops carry a cycle-count watchdog that traps instead of hanging the GPU.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace amira {
namespace {

using namespace tc;

constexpr int W_SL = 64;                                    // output features per slice (UMMA N)
constexpr int W_BM = 128;                                   // streams per M-tile (UMMA M)
constexpr int W_NG = kG / W_SL;                             // 40 slices per gate matrix
constexpr int W_NC = kH / W_SL;                             // 10
constexpr int W_ND = (kV + W_SL - 1) / W_SL;                // 17
constexpr int W_CTAS = 3 * W_NG + W_NC + W_ND;              // 147
constexpr int W_ND2 = W_ND + 1;                             // cluster variant: vocabulary padded to an even slice count
constexpr int W_CTAS2 = 3 * W_NG + W_NC + W_ND2;            // 148 = every SM of a B200, 74 CTA pairs
constexpr int W_KC = kH / BK;                               // 10 k-chunks
constexpr int W_WCHUNK = W_SL * BK * 2;                     // 8 KB: [64 rows][64 k] bf16
constexpr int W_WBYTES = 2 * W_KC * W_WCHUNK;               // 160 KB: per k-chunk [w_hi rows 0-63 | w_lo rows 64-127]
constexpr int W_UNIT = W_BM * BK * 2;                       // 16 KB: [128 rows][64 k] bf16 — also exactly one fused weight k-chunk
constexpr int W_RING = 4;
// Weight k-chunks kept resident (of 10).  With W_RES < 10 the other chunks' 16 KB stream through the ring with the
// activations and every chunk given up adds 16 KB of ring.  Measured (A/B on one GPU): W_RES = 7 (112 KB ring) is 3.7 %
// SLOWER than 10 (64 KB ring) — the unit is bound by the shared-memory port (TMA writes + operand reads), not by the bytes
// in flight, so the extra 15 % of bytes per unit costs more than the deeper ring hides.
constexpr int W_RES = 10;
constexpr int W_RING_MAX = W_RING + W_KC;
constexpr int W_CTRL = 2048;
constexpr int W_SMEM = W_WBYTES + W_RING * W_UNIT + W_CTRL;  // 231424 of the 232448-byte per-CTA maximum
constexpr int W_ACC_COLS = 2 * W_SL;                        // accumulator = [a_hi w_hi + a_lo w_hi | a_hi w_lo], summed in the epilogue
constexpr int W_NACC = 4;                                   // TMEM accumulators of 128 columns
constexpr int W_Q = 4;                                      // descriptor queue depth (= signal slots: bounds the epilogue's run-ahead)
constexpr int W_EPI_WARPS = 8, W_EPI_THREADS = W_EPI_WARPS * 32;
constexpr int W_THREADS = (4 + W_EPI_WARPS) * 32;           // 384: warp 0 TMA, 1 MMA, 2 scheduler, 3 idle, 4..11 epilogue
constexpr int W_MAX_MT = 256;
constexpr long long W_SPIN_LIMIT = 6000000000LL;            // ~3 s of SM clocks
constexpr int W_TRACE_ITS = 512;

enum { R_A = 0, R_BI = 1, R_BH = 2, R_C = 3, R_D = 4 };

struct WCtl {
    int t, sym, total, last, active, nsteps, failed, pad;
};

struct WsParams {
    CUtensorMap h0_hi, h0_lo, h1_hi, h1_lo, z_hi, z_lo;   // activations, box {64 k, 128 rows} (64 rows in the cluster variant)
    CUtensorMap whh0_hi, whh0_lo, w1_hi, w1_lo, wp_hi, wp_lo, wo_hi, wo_lo;                // weights, box {64, 64}
    const __nv_bfloat16 *g_whh0_hi, *g_whh0_lo, *g_w1_hi, *g_w1_lo, *g_wp_hi, *g_wp_lo, *g_wo_hi, *g_wo_lo;  // TS variant: weights -> TMEM
    const float *g0p, *b1p, *boutp, *E;
    int B, Mpad, MT, T;
    const int *lens, *slots, *perm, *eoff;
    __nv_bfloat16 *h0b_hi, *h0b_lo, *h1b_hi, *h1b_lo, *zb_hi, *zb_lo;
    float *h0f, *h1f, *c0, *c1;
    float *part;        // [MT][40 slices][2 column groups][128 rows][32] fp32: h1(t-1) W_hh1 partial sums
    unsigned long long *amax;  // [2 step parities][Mpad] packed (orderable logit << 32 | ~column): atomicMax = first-max argmax
    WCtl *ctl;
    int4 *rowinfo;      // [Mpad] {stream index, encoded length, first packed row of E, 0}: one load instead of perm -> lens / eoff chains
    int *tile_active, *cnt_d, *cnt_a, *cnt_b, *cnt_c, *dead_at, *part_ready /* [MT][40] */, *fail_count;
    float *s1, *s2;
    int *tokens, *ntok, *nsteps;
    int max_sym, max_total, blank, relu;
    int norot;          // debug: all CTAs walk the k-chunks in the same order
    int nrows;          // stream rows a unit loads and multiplies (TS form: 32 / 64 when a single M-tile holds that few streams, else 128)
    int variant;        // debug: AMIRA_WS_VARIANT bit mask of experimental code paths (A/B timing)
    int trace_role;     // debug: role whose slice-0 CTA is traced for every M-tile
    long long *trace;   // nullable: [W_TRACE_ITS][32] globaltimer stamps of M-tile 0 (debug)
};

struct WsDesc {
    int mt, it;
};
struct WsSmem {
    uint64_t full[W_RING_MAX], empty[W_RING_MAX], acc_full[W_NACC], acc_empty[W_NACC], q_full[W_Q], q_empty[W_Q], wfull;
    uint32_t tmem_slot;
    int act2[2];  // live-stream count of the unit in the layer-0 epilogue (alternating slots)
    uint64_t sig_full[W_Q];  // TS: all epilogue threads have issued the unit's stores -> the signal thread fences and publishes
    int sig_skip[W_Q];       // TS: the unit has nothing to publish (layer-0 unit of an M-tile that just ended)
    WsDesc q[W_Q];
    unsigned char dead[W_MAX_MT];
};
static_assert(sizeof(WsSmem) <= W_CTRL, "control block exceeds its reservation");
static_assert(10 * W_UNIT + W_BM * (W_BM + 1) * 4 + (int)sizeof(WsSmem) <= W_SMEM, "TS layout exceeds the shared-memory budget");

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// activations: exp-based, absolute error ~1e-7 (fp32 rounding level), ex2/rcp on the SFU
__device__ __forceinline__ float fsig(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// low word of the K-major SWIZZLE_128B shared-memory descriptor (address >> 4 | LBO 1); the high word is constant
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
    // high word: SBO = 1024 B >> 4 in [32,46), version 1 at bit 46, layout type 2 (128B swizzle) in [61,64)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(64u | (1u << 14) | (2u << 29))
        : "memory");
}
// ---- thread-block-cluster helpers (CTA pairs share every activation tile: each CTA loads half, TMA multicast to both) ----
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_s32(uint32_t cluster_addr, int v) {
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// arrive on the same mbarrier in every CTA of `mask` once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T: the A operand (row m in TMEM lane m, k-pair j in 32-bit column j) is read from tensor
// memory, so only B costs shared-memory bandwidth (measured 64 clk per M=128 N=128 K=16 instruction, scripts/bench_umma.cu)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(64u | (1u << 14) | (2u << 29))
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// debug trace: slice 0 of every role stamps its events for M-tile `W_TRACE_MT`
#define WS_TRACE(ev)                                                                                   \
    do {                                                                                               \
        if (p.trace && slice == 0 && it < W_TRACE_ITS) {                                               \
            if (mt == 0) p.trace[it * 32 + role * 6 + (ev)] = gtime();                                 \
            if (role == p.trace_role && mt < 8) p.trace[W_TRACE_ITS * 32 + (it * 8 + mt) * 8 + (ev)] = gtime(); \
        }                                                                                              \
    } while (0)
// AMIRA_WS_VARIANT bit 7: events 1, 2, 6, 7 of the per-tile trace are stamped by the epilogue (unit popped, head loads back,
// accumulator gathered, arithmetic done) instead of by the MMA thread
#define WS_TRACE_MMA(ev) do { if (!(p.variant & 128)) WS_TRACE(ev); } while (0)
#define WS_TRACE_EPI(ev) do { if ((p.variant & 128) && etid == 0) WS_TRACE(ev); } while (0)

// this thread's 32 accumulator columns: (a_hi w_hi + a_lo w_hi) + (a_hi w_lo), the two halves of the 128-column accumulator
__device__ __forceinline__ void tmem_ld32_sum(uint32_t taddr, uint32_t (&r)[32]) {
    tmem_ld32(taddr, r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t t[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]),
              "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
            : "r"(taddr + W_SL + h * 16)
            : "memory");
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) r[h * 16 + j] = __float_as_uint(__uint_as_float(r[h * 16 + j]) + __uint_as_float(t[j]));
    }
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_ge(const int *p, int target) {
    if (ld_acquire(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire(p) < target) {
        if (clock64() - t0 > W_SPIN_LIMIT) __trap();
    }
}
// wait until *cnt >= target (returns 0) or the M-tile is known to have ended before iteration `it` (returns 1)
__device__ __forceinline__ int spin_ge_or_dead_acq(const int *cnt, int target, const int *dead_at, int it, int lazy_dead) {
    const long long t0 = clock64();
    if (ld_acquire(dead_at) <= it) return 1;
    for (uint32_t n = 0;; ++n) {  // the end-of-tile marker is a rare event: with lazy_dead it is polled every 4th miss only, which
        if (ld_acquire(cnt) >= target) return 0;  // halves the poll period (one L2 round trip) and the detection delay with it
        if ((!lazy_dead || (n & 3) == 3) && ld_acquire(dead_at) <= it) return 1;
        if (clock64() - t0 > W_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ int spin_ge_or_dead(const int *cnt, int target, const int *dead_at, int it) {
    const long long t0 = clock64();
    for (;;) {  // both polls in flight together (relaxed), one acquire fence on the way out
        const int d = ld_relaxed(dead_at), v = ld_relaxed(cnt);
        if (d <= it) { fence_acq_rel_gpu(); return 1; }
        if (v >= target) { fence_acq_rel_gpu(); return 0; }
        if (clock64() - t0 > W_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++n & 0xfff) == 0 && clock64() - t0 > W_SPIN_LIMIT) __trap();
    }
}
// a queue consumer is done with a descriptor slot: rank 0 owns the queue barriers
template <int CL>
__device__ __forceinline__ void q_release(uint64_t *q_empty_local, uint32_t crank) {
    if (CL == 1 || crank == 0) mbar_arrive(q_empty_local);
    else mbar_arrive_cluster(map_to_cta(smem_u32(q_empty_local), 0));
}
__device__ __forceinline__ WCtl load_ctl(const WCtl *q) {  // L1-bypassing: written by another SM's control update
    const int4 a = __ldcg(reinterpret_cast<const int4 *>(q)), b = __ldcg(reinterpret_cast<const int4 *>(q) + 1);
    WCtl c;
    c.t = a.x; c.sym = a.y; c.total = a.z; c.last = a.w; c.active = b.x; c.nsteps = b.y; c.failed = b.z; c.pad = b.w;
    return c;
}
__device__ __forceinline__ size_t ws_state_off(const WsParams &p, int layer, int b) {
    return p.slots ? ((size_t)p.slots[b] * 2 + layer) * kH : ((size_t)layer * p.B + b) * kH;
}

// CL = 1: 147 independent CTAs.  CL = 2: 148 CTAs in 74 clusters of two neighbouring slices of one role; the pair shares every
// activation tile (each CTA loads 64 of the 128 rows, TMA multicast delivers them to both), which halves the L2 read traffic
// that bounds the kernel while many M-tiles are in flight.  Rank 0's scheduler drives both CTAs in lock step.
// TS = true ("tensor-memory-stationary"): the slice's [w_hi ; w_lo] lives in 320 columns of TENSOR memory as the A operand
// (M = 128 rows: 64 hi + 64 lo), the activation tile is the B operand (N = 128 streams) and the accumulator comes out transposed
// (feature parts along the lanes, streams along the columns).  Shared memory then holds no weights at all: the TMA ring grows
// from 4 to 9 slots, the tcgen05 operand reads drop from 14 KB to 8 KB per k-step (the shared-memory port is what bounds a
// unit), and the epilogue transposes the accumulator through a 66 KB shared tile back to the stream-per-thread layout.
template <int CL, bool TS>
__global__ void __launch_bounds__(W_THREADS, 1) greedy_ws_kernel(const __grid_constant__ WsParams p) {
    cg::grid_group grid = cg::this_grid();
    static_assert(!(TS && CL != 1), "the tensor-memory-stationary variant is single-CTA");
    constexpr int ND = CL == 2 ? W_ND2 : W_ND;
    constexpr int NACC = TS ? 1 : W_NACC;              // TS: one 128-column accumulator behind the 320 weight columns
    constexpr uint32_t ACC_COL0 = TS ? 320 : 0;
    constexpr int T_LD = W_BM + 1;                     // TS: row stride (floats) of the transposition tile
    const uint32_t crank = CL == 2 ? (blockIdx.x & 1) : 0;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int RES = TS ? 0 : (CL == 1 ? W_RES : W_KC);  // weight k-chunks resident in SHARED memory (CTA pairs share the ring: all)
    constexpr int NRING = TS ? 10 : W_RING + (W_KC - RES);  // ring slots of 16 KB; RES + NRING = 14 slots = 224 KB unless TS
    unsigned char *w_hi = smem, *ring = smem + RES * 2 * W_WCHUNK;  // w_hi: [resident k-chunk][w_hi 8 KB | w_lo 8 KB]
    float *ttile = reinterpret_cast<float *>(smem + 10 * W_UNIT);  // TS: [128 feature parts][T_LD] accumulator transposition tile
    // control block: behind the 224 KB of weights + ring, or (TS) behind the 160 KB ring and the 64.5 KB transposition tile
    WsSmem &sm = *reinterpret_cast<WsSmem *>(smem + (TS ? 10 * W_UNIT + W_BM * T_LD * 4 : W_WBYTES + W_RING * W_UNIT));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // role and slice of this CTA
    int role, slice;
    {
        const int b = blockIdx.x;
        if (b < W_NG) { role = R_A; slice = b; }
        else if (b < 2 * W_NG) { role = R_BI; slice = b - W_NG; }
        else if (b < 3 * W_NG) { role = R_BH; slice = b - 2 * W_NG; }
        else if (b < 3 * W_NG + W_NC) { role = R_C; slice = b - 3 * W_NG; }
        else { role = R_D; slice = b - 3 * W_NG - W_NC; }
    }

    // CTAs of one phase read the same activation tile at the same time: each starts at a different k-chunk so the requests
    // spread over the tile's L2 slices instead of queueing on one 16 KB region
    const int kc0 = p.norot ? 0 : ((slice / CL) * 3 + role) % W_KC;  // one order per cluster: its CTAs share the ring contents

    if (tid == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled operand layout needs a 1024-byte aligned base
        for (int s = 0; s < NRING; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], CL); }  // empty: every CTA's MMA commit
        for (int b = 0; b < NACC; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], W_EPI_THREADS); }
        // queue consumers: TMA thread, MMA thread, one lane per epilogue warp — of every CTA of the cluster (rank 0 owns the queue)
        for (int i = 0; i < W_Q; ++i) {
            mbar_init(&sm.q_full[i], 1);
            mbar_init(&sm.q_empty[i], CL * (2 + W_EPI_WARPS) + (TS ? 1 : 0));  // + the signal thread
            mbar_init(&sm.sig_full[i], W_EPI_THREADS);
            sm.sig_skip[i] = 0;
        }
        mbar_init(&sm.wfull, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < W_MAX_MT; i += W_THREADS) sm.dead[i] = 0;
    if (tid == 0) { sm.act2[0] = 0; sm.act2[1] = 0; }
    if (warp == 1) tmem_alloc(&sm.tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (CL == 2) cluster_sync_all();  // the peer's barriers exist before any multicast copy or remote arrive targets them

    // ---- stationary weights: one TMA burst, overlapped with the prologue below ----
    if (!TS && warp == 0 && lane == 0) {
        const CUtensorMap *mh, *ml;
        int kcol = 0;
        if (role == R_A) { mh = &p.whh0_hi; ml = &p.whh0_lo; }
        else if (role == R_BI) { mh = &p.w1_hi; ml = &p.w1_lo; }
        else if (role == R_BH) { mh = &p.w1_hi; ml = &p.w1_lo; kcol = kH; }
        else if (role == R_C) { mh = &p.wp_hi; ml = &p.wp_lo; }
        else { mh = &p.wo_hi; ml = &p.wo_lo; }
        tma_prefetch_desc(mh);
        tma_prefetch_desc(ml);
        mbar_expect_tx(&sm.wfull, RES * 2 * W_WCHUNK);
        for (int kc = 0; kc < RES; ++kc) {  // w_lo directly below w_hi: together one 128-row B operand
            tma_load_2d(w_hi + kc * 2 * W_WCHUNK, mh, &sm.wfull, kcol + kc * BK, slice * W_SL);
            tma_load_2d(w_hi + kc * 2 * W_WCHUNK + W_WCHUNK, ml, &sm.wfull, kcol + kc * BK, slice * W_SL);
        }
    }

    if (TS && warp >= 4) {  // weights -> tensor memory: TMEM lane L = A row (L < 64: w_hi of feature L, else w_lo of feature L - 64)
        const int q_ = warp & 3, ch = (warp - 4) >> 2, L = q_ * 32 + lane, part = L >> 6, j = L & 63;
        const __nv_bfloat16 *gh, *gl;
        int ldw = kH, kcol = 0;
        if (role == R_A) { gh = p.g_whh0_hi; gl = p.g_whh0_lo; }
        else if (role == R_BI) { gh = p.g_w1_hi; gl = p.g_w1_lo; ldw = 2 * kH; }
        else if (role == R_BH) { gh = p.g_w1_hi; gl = p.g_w1_lo; ldw = 2 * kH; kcol = kH; }
        else if (role == R_C) { gh = p.g_wp_hi; gl = p.g_wp_lo; }
        else { gh = p.g_wo_hi; gl = p.g_wo_lo; }
        const uint4 *src = reinterpret_cast<const uint4 *>((part ? gl : gh) + (size_t)(slice * W_SL + j) * ldw + kcol + ch * (kH / 2));
#pragma unroll 1
        for (int blk = 0; blk < 5; ++blk) {  // 5 x 32 columns = 320 bf16 = this thread's half of the row
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 t4 = __ldg(src + blk * 8 + i);
                v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
            }
            tmem_st32(sm.tmem_slot + ((uint32_t)(q_ * 32) << 16) + ch * (kH / 4) + blk * 32, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
    }

    // ---- prologue: initial LSTM state (fp32 + split bf16, parity 0), control, default results ----
    const size_t n_state = (size_t)p.Mpad * kH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n_state; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / kH), j = (int)(i % kH);
        const int b = row < p.B ? p.perm[row] : -1;
        const float h0 = (b >= 0 && p.s1) ? p.s1[ws_state_off(p, 0, b) + j] : 0.f;
        const float h1 = (b >= 0 && p.s1) ? p.s1[ws_state_off(p, 1, b) + j] : 0.f;
        p.h0f[i] = h0;
        p.h1f[i] = h1;
        p.c0[i] = (b >= 0 && p.s2) ? p.s2[ws_state_off(p, 0, b) + j] : 0.f;
        p.c1[i] = (b >= 0 && p.s2) ? p.s2[ws_state_off(p, 1, b) + j] : 0.f;
        __nv_bfloat16 hh, hl;
        split_bf16(h0, hh, hl);
        p.h0b_hi[i] = hh; p.h0b_lo[i] = hl;
        split_bf16(h1, hh, hl);
        p.h1b_hi[i] = hh; p.h1b_lo[i] = hl;
    }
    for (int row = blockIdx.x * blockDim.x + tid; row < p.Mpad; row += gridDim.x * blockDim.x) {
        WCtl c;
        c.t = 0; c.sym = 0; c.total = 0; c.last = p.blank; c.nsteps = 0; c.failed = 0; c.pad = 0;
        c.active = (row < p.B && p.lens[p.perm[row]] > 0) ? 1 : 0;
        p.ctl[row] = c;  // the control rows are double-buffered by step parity (see the layer-0 epilogue)
        p.ctl[p.Mpad + row] = c;
        const int prow_ = row < p.B ? p.perm[row] : 0;
        p.rowinfo[row] = make_int4(prow_, row < p.B ? p.lens[prow_] : 0, row < p.B ? p.eoff[prow_] : 0, 0);
        if (c.active) atomicAdd(&p.tile_active[row / W_BM], 1);
        if (row < p.B) {
            p.ntok[p.perm[row]] = 0;
            if (p.nsteps) p.nsteps[p.perm[row]] = 0;
        }
    }
    __threadfence();
    fence_proxy_async();
    grid.sync();
    // M-tiles with no active stream never start (dead_at = 0); the host zero-initialised every counter
    for (int mt = blockIdx.x * blockDim.x + tid; mt < p.MT; mt += gridDim.x * blockDim.x)
        p.dead_at[mt] = __ldcg(p.tile_active + mt) > 0 ? 0x7fffffff : 0;
    __threadfence();
    grid.sync();

    const uint32_t tmem_base = sm.tmem_slot;

    if (warp == 2) {
        if (lane == 0 && crank == 0) {  // ===================== scheduler: polls the dependency counters, publishes runnable units =====================
            // A thread of its own so that the TMA thread never stalls on a global-memory poll: the next unit's loads go out
            // the moment ring slots free up, and the first-load latency hides behind the tail of the current unit.
            uint32_t qn = 0;
            auto publish = [&](int mt_, int it_, uint32_t &n) {  // into the queue of every CTA of the cluster
                const uint32_t slot = n % W_Q;
                mbar_wait_wd(&sm.q_empty[slot], ((n / W_Q) & 1) ^ 1);
                sm.q[slot].mt = mt_; sm.q[slot].it = it_;
                mbar_arrive(&sm.q_full[slot]);
                if (CL == 2) {
                    st_cluster_s32(map_to_cta(smem_u32(&sm.q[slot].mt), 1), mt_);
                    st_cluster_s32(map_to_cta(smem_u32(&sm.q[slot].it), 1), it_);
                    mbar_arrive_cluster(map_to_cta(smem_u32(&sm.q_full[slot]), 1));  // release.cluster orders the two stores
                }
                ++n;
            };
            for (int it = 0;; ++it) {
                bool any = false;
                for (int mt = 0; mt < p.MT; ++mt) {
                    if (sm.dead[mt]) continue;
                    int st;
                    if (p.variant & 1) {
                        const int *cp = (role == R_A || role == R_BI) ? p.cnt_a + mt : (role == R_D ? p.cnt_c + mt : p.cnt_b + mt);
                        const int tg = (role == R_A ? W_NG * it : (role == R_BI ? W_NG * (it + 1) : (role == R_BH ? W_NG * it : (role == R_C ? W_NG * (it + 1) : W_NC * (it + 1)))));
                        st = spin_ge_or_dead_acq(cp, tg, p.dead_at + mt, it, p.variant & 8);
                    } else
                    if (role == R_A) st = spin_ge_or_dead(p.cnt_a + mt, W_NG * it, p.dead_at + mt, it);              // h0(it-1)
                    else if (role == R_BI) st = spin_ge_or_dead(p.cnt_a + mt, W_NG * (it + 1), p.dead_at + mt, it);  // h0(it)
                    else if (role == R_BH) st = spin_ge_or_dead(p.cnt_b + mt, W_NG * it, p.dead_at + mt, it);        // h1(it-1)
                    else if (role == R_C) st = spin_ge_or_dead(p.cnt_b + mt, W_NG * (it + 1), p.dead_at + mt, it);   // h1(it)
                    else st = spin_ge_or_dead(p.cnt_c + mt, W_NC * (it + 1), p.dead_at + mt, it);                    // z(it)
                    if (st) { sm.dead[mt] = 1; continue; }
                    any = true;
                    WS_TRACE(0);
                    publish(mt, it, qn);
                }
                if (!any) break;
            }
            publish(-1, 0, qn);  // exit descriptor
        }
    } else if (warp == 0) {
        if (lane == 0) {  // ===================== TMA producer =====================
            const CUtensorMap *a_hi, *a_lo;
            if (role == R_A || role == R_BI) { a_hi = &p.h0_hi; a_lo = &p.h0_lo; }
            else if (role == R_BH || role == R_C) { a_hi = &p.h1_hi; a_lo = &p.h1_lo; }
            else { a_hi = &p.z_hi; a_lo = &p.z_lo; }
            tma_prefetch_desc(a_hi);
            tma_prefetch_desc(a_lo);
            const CUtensorMap *wmh, *wml;  // this slice's weights, for the k-chunks that are not resident
            int wcol = 0;
            if (role == R_A) { wmh = &p.whh0_hi; wml = &p.whh0_lo; }
            else if (role == R_BI) { wmh = &p.w1_hi; wml = &p.w1_lo; }
            else if (role == R_BH) { wmh = &p.w1_hi; wml = &p.w1_lo; wcol = kH; }
            else if (role == R_C) { wmh = &p.wp_hi; wml = &p.wp_lo; }
            else { wmh = &p.wo_hi; wml = &p.wo_lo; }
            uint32_t u = 0, qn = 0;
            for (;;) {
                const uint32_t slot = qn % W_Q;
                mbar_wait_wd(&sm.q_full[slot], (qn / W_Q) & 1);
                const int mt = sm.q[slot].mt, it = sm.q[slot].it;
                q_release<CL>(&sm.q_empty[slot], crank);
                ++qn;
                if (mt < 0) break;
                const int par = it & 1;
                // activations of step it-1 for the recurrent roles (parity par), of step it for the others (parity par^1)
                const int a_row = (role == R_D) ? 0 : ((role == R_A || role == R_BH) ? par : (par ^ 1)) * p.Mpad;
                fence_proxy_async();  // the scheduler's acquire (through the queue barrier) before these async-proxy reads
                for (int ki = 0; ki < W_KC; ++ki) {
                    const int kc = (kc0 + ki) % W_KC;
                    if (!TS && kc >= RES) {  // this chunk's weights travel with the activations: one slot, [w_hi ; w_lo] like the resident ones
                        const uint32_t s = u % NRING;
                        mbar_wait_wd(&sm.empty[s], ((u / NRING) & 1) ^ 1);
                        mbar_expect_tx(&sm.full[s], W_UNIT);
                        tma_load_2d(ring + s * W_UNIT, wmh, &sm.full[s], wcol + kc * BK, slice * W_SL);
                        tma_load_2d(ring + s * W_UNIT + W_WCHUNK, wml, &sm.full[s], wcol + kc * BK, slice * W_SL);
                        ++u;
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t s = u % NRING;
                        mbar_wait_wd(&sm.empty[s], ((u / NRING) & 1) ^ 1);
                        mbar_expect_tx(&sm.full[s], TS ? (uint32_t)p.nrows * (BK * 2) : (uint32_t)W_UNIT);
                        if (CL == 1) tma_load_2d(ring + s * W_UNIT, half ? a_lo : a_hi, &sm.full[s], kc * BK, a_row + mt * W_BM);
                        else  // this CTA's 64 rows of the tile, delivered to both CTAs (each full barrier sees 2 x 8 KB)
                            tma_load_2d_mc(ring + s * W_UNIT + crank * (W_UNIT / 2), half ? a_lo : a_hi, &sm.full[s], kc * BK,
                                           a_row + mt * W_BM + (int)crank * (W_BM / 2), (uint16_t)3);
                        ++u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===================== MMA issuer =====================
            // descriptors are (smem address >> 4) in the low word plus constant fields: precomputed so the issue loop is a handful
            // of integer adds per tcgen05.mma (a single thread issuing ~100 of them per unit must not be the bottleneck)
            constexpr uint32_t idesc_cat = make_idesc_bf16(W_BM, 2 * W_SL), idesc_hi = make_idesc_bf16(W_BM, W_SL);
            const uint32_t w_lo32 = sdesc_lo(smem_u32(w_hi)), ring_lo32 = sdesc_lo(smem_u32(ring));
            if (!TS) mbar_wait_wd(&sm.wfull, 0);
            uint32_t u = 0, qn = 0, tile = 0;
            for (;;) {
                const uint32_t slot = qn % W_Q;
                mbar_wait_wd(&sm.q_full[slot], (qn / W_Q) & 1);
                const int mt = sm.q[slot].mt, it = sm.q[slot].it;
                q_release<CL>(&sm.q_empty[slot], crank);
                ++qn;
                if (mt < 0) break;
                const uint32_t buf = tile % NACC, use = tile / NACC;
                mbar_wait_wd(&sm.acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + ACC_COL0 + buf * W_ACC_COLS;
                int kc = kc0;
                long long wait_cyc = 0;
                if (TS) {
                    // A = the slice's [w_hi ; w_lo] in tensor memory (32 columns per k-chunk), B = activation tile (N = 128 streams):
                    // two instructions per k-step, a_hi then a_lo, both against all 128 weight rows (w_lo x a_lo is a free 2^-18 term).
                    // One thread issues all of it, and a lone thread retires an instruction only every few cycles: measured
                    // (scripts/bench_pipe.cu) ~150 instructions of bookkeeping per ring slot made the ISSUE loop, not the tensor pipe
                    // or shared memory, the limit of a unit.  Hence: 20 slots per unit on a 10-slot ring, so slot index and barrier
                    // parity are compile-time constants of the fully unrolled loop; descriptors are one add; no watchdog here.
                    static_assert(!TS || NRING == 10, "slot / parity constants below assume two ring revolutions per unit");
                    const uint32_t idesc_t = make_idesc_bf16(W_BM, p.nrows);  // M = 128 weight rows (hi | lo), N = stream rows
#pragma unroll
                    for (int ki = 0; ki < W_KC; ++ki) {
                        const uint32_t wa = tmem_base + kc * 32;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int j = 2 * ki + half, s = j % 10;
                            mbar_wait(&sm.full[s], (uint32_t)(j / 10));
                            if (j == 0) WS_TRACE_MMA(1);
                            const uint32_t bd = ring_lo32 + s * (W_UNIT >> 4);
                            umma_bf16_ts(acc, wa, bd, idesc_t, j != 0);
                            umma_bf16_ts(acc, wa + 8, bd + 2, idesc_t, 1);
                            umma_bf16_ts(acc, wa + 16, bd + 4, idesc_t, 1);
                            umma_bf16_ts(acc, wa + 24, bd + 6, idesc_t, 1);
                            umma_commit(&sm.empty[s]);
                        }
                        kc = kc + 1 == W_KC ? 0 : kc + 1;
                    }
                    u += 2 * W_KC;
                } else {
#pragma unroll 1
                for (int ki = 0; ki < W_KC; ++ki) {
                    uint32_t wd = w_lo32 + kc * (2 * W_WCHUNK >> 4);
                    uint32_t ws = 0xffffffffu;  // ring slot of a streamed weight chunk
                    if (kc >= RES) {
                        ws = u % NRING;
                        mbar_wait_wd(&sm.full[ws], (u / NRING) & 1);
                        wd = ring_lo32 + ws * (W_UNIT >> 4);
                        ++u;
                    }
                    {   // hi unit: a_hi * [w_hi ; w_lo]  (N = 128)
                        const uint32_t s = u % NRING;
                        const long long tw0 = p.trace ? clock64() : 0;
                        mbar_wait_wd(&sm.full[s], (u / NRING) & 1);
                        if (p.trace && ki > 0) wait_cyc += clock64() - tw0;
                        if (ki == 0) WS_TRACE_MMA(1);
                        if (p.variant & 16) tc_fence_after();
                        const uint32_t ad = ring_lo32 + s * (W_UNIT >> 4);
                        umma_bf16_lo(acc, ad, wd, idesc_cat, ki != 0);
                        umma_bf16_lo(acc, ad + 2, wd + 2, idesc_cat, 1);
                        umma_bf16_lo(acc, ad + 4, wd + 4, idesc_cat, 1);
                        umma_bf16_lo(acc, ad + 6, wd + 6, idesc_cat, 1);
                        if (CL == 1) umma_commit(&sm.empty[s]); else umma_commit_mc(&sm.empty[s], (uint16_t)3);  // slot free in both CTAs
                        ++u;
                    }
                    {   // lo unit: a_lo * w_hi  (N = 64, accumulator columns 0..63)
                        const uint32_t s = u % NRING;
                        const long long tw0 = p.trace ? clock64() : 0;
                        mbar_wait_wd(&sm.full[s], (u / NRING) & 1);
                        if (p.trace) wait_cyc += clock64() - tw0;
                        if (p.variant & 16) tc_fence_after();
                        const uint32_t ad = ring_lo32 + s * (W_UNIT >> 4);
                        umma_bf16_lo(acc, ad, wd, idesc_hi, 1);
                        umma_bf16_lo(acc, ad + 2, wd + 2, idesc_hi, 1);
                        umma_bf16_lo(acc, ad + 4, wd + 4, idesc_hi, 1);
                        umma_bf16_lo(acc, ad + 6, wd + 6, idesc_hi, 1);
                        if (CL == 1) umma_commit(&sm.empty[s]); else umma_commit_mc(&sm.empty[s], (uint16_t)3);  // slot free in both CTAs
                        if (ws != 0xffffffffu) umma_commit(&sm.empty[ws]);  // both products of the chunk have read the streamed weights
                        ++u;
                    }
                    kc = kc + 1 == W_KC ? 0 : kc + 1;
                }
                }
                umma_commit(&sm.acc_full[buf]);
                WS_TRACE_MMA(2);
                if (p.trace && role == R_D && slice == 0 && mt == 0 && it < W_TRACE_ITS) p.trace[it * 32 + 31] = wait_cyc;
                ++tile;
            }
        }
    } else if (TS && warp == 3) {
        if (lane == 0) {  // ===================== signal thread (TS) =====================
            // Publishing a unit costs a gpu-scope fence that waits for the CTA's stores to be acknowledged (1 us idle, 3 us under
            // load).  The epilogue threads only ARRIVE on a shared-memory barrier once their stores are issued and go on to the next
            // unit; this thread waits for the 256 arrivals, fences (cumulative over the stores it observed through the barrier) and
            // bumps the dependency counter.  The descriptor slot is released last, so the epilogue is never more than W_Q units ahead.
            uint32_t qn = 0;
            for (;;) {
                const uint32_t slot = qn % W_Q;
                mbar_wait_wd(&sm.q_full[slot], (qn / W_Q) & 1);
                const int mt = sm.q[slot].mt, it = sm.q[slot].it;
                if (mt < 0) { q_release<CL>(&sm.q_empty[slot], crank); break; }
                mbar_wait_wd(&sm.sig_full[slot], (qn / W_Q) & 1);
                const int skip = sm.sig_skip[slot];
                sm.sig_skip[slot] = 0;
                if (!skip) {
                    if (!(p.variant & 128)) WS_TRACE(6);  // every epilogue thread has arrived
                    __threadfence();  // (fence.acq_rel.gpu instead was measured identical)
                    if (!(p.variant & 128)) WS_TRACE(7);  // their stores are visible device-wide
                    if (role == R_A) { fence_proxy_async(); atomicAdd(p.cnt_a + mt, 1); }
                    else if (role == R_BI) { fence_proxy_async(); atomicAdd(p.cnt_b + mt, 1); }
                    else if (role == R_BH) st_release(p.part_ready + mt * W_NG + slice, it + 1);
                    else if (role == R_C) { fence_proxy_async(); atomicAdd(p.cnt_c + mt, 1); }
                    else atomicAdd(p.cnt_d + mt, 1);
                    WS_TRACE(4);
                }
                q_release<CL>(&sm.q_empty[slot], crank);
                ++qn;
            }
        }
    } else if (warp >= 4) {  // ===================== epilogue: 8 warps =====================
        const int e = warp - 4, q = warp & 3, cgp = e >> 2;  // TMEM lane quarter (must be warp % 4), 32-column group
        const int etid = tid - 128;
        const int r_in = q * 32 + lane;
        const int nb = slice * W_SL + cgp * 32;           // first of this thread's 32 output columns
        // this thread's 32 pre-activations (stream r_in, features nb .. nb+31) out of the accumulator, which is then handed back
        // TS, step 1: drain this thread's TMEM lane (a feature part) for 64 of the 128 streams into the shared transposition tile and
        // hand the accumulator back to the MMA thread (the tile is the second buffer that lets the next unit's MMAs run meanwhile)
        auto drain_acc = [&](uint32_t taddr, uint32_t buf, uint32_t (&r)[32]) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                tmem_ld32(taddr + h * 32, r);
                tmem_ld_wait();
                float *dstt = ttile + r_in * T_LD + cgp * 64 + h * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) dstt[j] = __uint_as_float(r[j]);
            }
            tc_fence_before();
            mbar_arrive(&sm.acc_empty[buf]);
            named_bar_sync(2, W_EPI_THREADS);
        };
        // TS, step 2: gather this thread's stream column — hi-part row + lo-part row of each of its 32 features
        auto gather_acc = [&](uint32_t (&r)[32]) {
            const float *srct = ttile + (cgp * 32) * T_LD + r_in;
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(srct[j * T_LD] + srct[(W_SL + j) * T_LD]);
            named_bar_sync(2, W_EPI_THREADS);  // the tile is free for the next unit
        };
        // this thread's 32 pre-activations (stream r_in, features nb .. nb+31) out of the accumulator, which is then handed back
        auto load_acc = [&](uint32_t taddr, uint32_t buf, uint32_t (&r)[32]) {
            if (!TS) {
                tmem_ld32_sum(taddr, r);
                tc_fence_before();
                mbar_arrive(&sm.acc_empty[buf]);
            } else {
                drain_acc(taddr, buf, r);
                gather_acc(r);
            }
        };
        // the same, summed in place into the 32 addends the caller already holds (bias / gathered row / partial sums): holding
        // accumulator and addends as separate arrays costs 64 registers at the 168-register cap and spilled on the step chain
        auto gather_add = [&](float *pre) {
            const float *srct = ttile + (cgp * 32) * T_LD + r_in;
#pragma unroll
            for (int j = 0; j < 32; ++j) pre[j] = (srct[j * T_LD] + srct[(W_SL + j) * T_LD]) + pre[j];
            named_bar_sync(2, W_EPI_THREADS);  // the tile is free for the next unit
        };
        auto load_acc_add = [&](uint32_t taddr, uint32_t buf, float *pre) {
            uint32_t r[32];
            if (!TS) {
                tmem_ld32_sum(taddr, r);
                tc_fence_before();
                mbar_arrive(&sm.acc_empty[buf]);
#pragma unroll
                for (int j = 0; j < 32; ++j) pre[j] = __uint_as_float(r[j]) + pre[j];
            } else {
                drain_acc(taddr, buf, r);
                gather_add(pre);
            }
        };
        uint32_t qn = 0, tile = 0;
        for (;;) {
            const uint32_t slot = qn % W_Q;
            mbar_wait_wd(&sm.q_full[slot], (qn / W_Q) & 1);
            const WsDesc d = sm.q[slot];
            __syncwarp();
            if (lane == 0) q_release<CL>(&sm.q_empty[slot], crank);
            ++qn;
            if (d.mt < 0) break;
            const int mt = d.mt, it = d.it, par = it & 1;
            const uint32_t buf = tile % NACC, use = tile / NACC;
            ++tile;
            const int row = mt * W_BM + r_in;
            const uint32_t taddr = tmem_base + ACC_COL0 + buf * W_ACC_COLS + ((uint32_t)(q * 32) << 16) + cgp * (TS ? 64 : 32);
            uint32_t r[32];

            if (role == R_A) {
                // The recurrent GEMM ran ahead.  This epilogue first applies the control flow of the previous step
                // (decoder_optimized.rs:133-188) to its 128 streams: the vocabulary CTAs left each stream's first-max argmax as
                // one packed 64-bit key (atomicMax), so every layer-0 CTA derives the same control rows redundantly from one
                // 8-byte load per stream and no separate control phase sits on the critical path.  Slice 0 publishes the rows
                // (double-buffered by step parity: other layer-0 CTAs may still be reading the previous ones), the tokens and,
                // when no stream of the M-tile is left, the results and the end marker.
                if (TS) {  // the recurrent GEMM ran ahead: park its result in the shared tile now, so the NEXT unit's GEMM can run
                    mbar_wait_wd(&sm.acc_full[buf], use & 1);  // while this epilogue waits for the previous step's vocabulary phase
                    tc_fence_after();
                    drain_acc(taddr, buf, r);
                }
                float4 ad[8], cold4[2];
                float *cst = p.c0 + (size_t)row * kH + nb / 4;
                cold4[0] = __ldcg(reinterpret_cast<const float4 *>(cst));  // loads that do not depend on the previous step's
                cold4[1] = __ldcg(reinterpret_cast<const float4 *>(cst) + 1);  // vocabulary phase go out before the wait on it
                const int4 ri = __ldg(p.rowinfo + row);
                WCtl c = load_ctl(p.ctl + (size_t)((it & 1) ^ 1) * p.Mpad + row);  // state after the update of step it-2
                if (it > 0) {
                    if (p.variant & 2) spin_ge(p.cnt_d + mt, ND * it);
                    else {
                        if (etid == 0) spin_ge(p.cnt_d + mt, ND * it);  // every vocabulary slice of step it-1 has merged its argmax
                        named_bar_sync(1, W_EPI_THREADS);
                    }
                }
                if (etid == 0) WS_TRACE(5);
                int &act_cnt = sm.act2[tile & 1];  // `tile` was advanced above: consecutive units alternate slots
                if (it > 0 && c.active) {
                    const unsigned long long key = __ldcg(p.amax + (size_t)((it - 1) & 1) * p.Mpad + row);
                    const int bi = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
                    const int len = ri.y;
                    c.nsteps += 1;                       // state carried unconditionally (decoder_optimized.rs:154)
                    c.sym += 1;                          // :133
                    if (bi == p.blank) {                 // :171-173
                        c.t += 1; c.sym = 0;
                        if (c.t >= len) c.active = 0;
                    } else {
                        if (slice == 0 && cgp == 0) p.tokens[(size_t)ri.x * p.max_total + c.total] = bi;   // :176
                        c.total += 1;
                        c.last = bi;
                        if (c.total >= p.max_total) c.active = 0;                    // :179-188
                        else if (c.sym >= p.max_sym) {                               // :133-137
                            c.t += 1; c.sym = 0;
                            if (c.t >= len) c.active = 0;
                        }
                        if (c.active && bi >= kEmbRows) { c.active = 0; c.failed = 1; }  // next step would fail (:148-152)
                    }
                }
                if (slice == 0 && cgp == 0) {  // the control rows of step `it`, for every other role's epilogue
                    int4 *dstc = reinterpret_cast<int4 *>(p.ctl + (size_t)(it & 1) * p.Mpad + row);
                    __stcg(dstc, make_int4(c.t, c.sym, c.total, c.last));
                    __stcg(dstc + 1, make_int4(c.active, c.nsteps, c.failed, 0));
                }
                if (cgp == 0 && c.active) atomicAdd(&act_cnt, 1);
                const bool act = c.active;
                if (act) {  // the token-dependent gather overlaps the barrier below
                    const float4 *addp = reinterpret_cast<const float4 *>(p.g0p + (size_t)c.last * kG + nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) ad[j] = __ldg(addp + j);
                }
                named_bar_sync(1, W_EPI_THREADS);
                const bool live = act_cnt > 0;  // identical in every layer-0 CTA
                if (etid == 0) sm.act2[(tile & 1) ^ 1] = 0;  // the next unit's slot: nobody touches it before that unit's barrier
                if (etid == 0 && slice == 0 && p.trace && mt == 0 && it > 0 && it - 1 < W_TRACE_ITS) p.trace[(it - 1) * 32 + 30] = gtime();
                if (!live && p.s1 && p.s2) {
                    // the M-tile is finished and every layer's state is final (the vocabulary phase of the last step was
                    // released by all of them): each layer-0 CTA hands back its 16 hidden features of the four state arrays
                    for (int i = etid; i < W_BM * (W_SL / 4); i += W_EPI_THREADS) {
                        const int grow = mt * W_BM + i / (W_SL / 4), j = slice * (W_SL / 4) + i % (W_SL / 4);
                        if (grow < p.B) {
                            const int b = __ldg(p.rowinfo + grow).x;
                            const size_t src = (size_t)grow * kH + j;
                            p.s1[ws_state_off(p, 0, b) + j] = __ldcg(p.h0f + src);
                            p.s1[ws_state_off(p, 1, b) + j] = __ldcg(p.h1f + src);
                            p.s2[ws_state_off(p, 0, b) + j] = __ldcg(p.c0 + src);
                            p.s2[ws_state_off(p, 1, b) + j] = __ldcg(p.c1 + src);
                        }
                    }
                }
                if (!live && slice == 0) {  // its streams' results, then the end marker
                    for (int rr = etid; rr < W_BM; rr += W_EPI_THREADS) {
                        const int grow = mt * W_BM + rr;
                        if (grow < p.B) {
                            const WCtl f = load_ctl(p.ctl + (size_t)(it & 1) * p.Mpad + grow);
                            const int b = __ldg(p.rowinfo + grow).x;
                            p.ntok[b] = f.failed ? -1 : f.total;
                            if (p.nsteps) p.nsteps[b] = f.nsteps;
                            if (f.failed) atomicAdd(p.fail_count, 1);
                        }
                    }
                    named_bar_sync(1, W_EPI_THREADS);
                    if (etid == 0) {
                        __threadfence();
                        st_release(p.dead_at + mt, it);  // iteration `it` of this M-tile does not exist
                    }
                }
                // pre-activations = accumulator + G0[token] row, summed in place in the gathered row's registers (holding both
                // as separate arrays spilled 24 registers to local memory on this on-chain epilogue)
                float *pre = reinterpret_cast<float *>(ad);
                if (TS) {
                    if (etid == 0) WS_TRACE(3);
                    gather_add(pre);  // drained into the shared tile at the top of this unit
                } else {
                    mbar_wait_wd(&sm.acc_full[buf], use & 1);
                    if (etid == 0) WS_TRACE(3);
                    tc_fence_after();
                    load_acc(taddr, buf, r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) pre[j] = __uint_as_float(r[j]) + pre[j];
                }
                if (!live) {  // speculative unit of an ended M-tile: drop it (uniform across the CTA)
                    if (TS) {
                        if (etid == 0) sm.sig_skip[(tile - 1) % W_Q] = 1;
                        mbar_arrive(&sm.sig_full[(tile - 1) % W_Q]);
                    }
                    continue;
                }
                if (act) {
                    float cold[8] = {cold4[0].x, cold4[0].y, cold4[0].z, cold4[0].w, cold4[1].x, cold4[1].y, cold4[1].z, cold4[1].w};
                    float hnew[8];
                    __align__(16) __nv_bfloat16 vh[8], vl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float gi = fsig(pre[4 * j + 0]), gf = fsig(pre[4 * j + 1]);
                        const float gg = ftanh(pre[4 * j + 2]), go = fsig(pre[4 * j + 3]);
                        const float cn = gf * cold[j] + gi * gg;
                        cold[j] = cn;
                        hnew[j] = go * ftanh(cn);
                        split_bf16(hnew[j], vh[j], vl[j]);
                    }
                    float *hf = p.h0f + (size_t)row * kH + nb / 4;
                    const size_t ob = ((size_t)(par ^ 1) * p.Mpad + row) * kH + nb / 4;
                    reinterpret_cast<float4 *>(cst)[0] = make_float4(cold[0], cold[1], cold[2], cold[3]);
                    reinterpret_cast<float4 *>(cst)[1] = make_float4(cold[4], cold[5], cold[6], cold[7]);
                    reinterpret_cast<float4 *>(hf)[0] = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
                    reinterpret_cast<float4 *>(hf)[1] = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
                    *reinterpret_cast<uint4 *>(p.h0b_hi + ob) = *reinterpret_cast<uint4 *>(vh);
                    *reinterpret_cast<uint4 *>(p.h0b_lo + ob) = *reinterpret_cast<uint4 *>(vl);
                }
                if (TS) {  // the unit's stores are issued: the signal thread fences and publishes them; this thread moves on
                    if (etid == 0) WS_TRACE(5);
                    mbar_arrive(&sm.sig_full[(tile - 1) % W_Q]);
                } else {
                    named_bar_sync(1, W_EPI_THREADS);
                    if (etid == 0) {  // cumulative release of every epilogue thread's stores (ordered by the barrier); the readers use TMA.
                        // (One release per thread instead — no barrier — was measured 3x slower: 10 k atomics per M-tile step on one word.)
                        __threadfence();
                        fence_proxy_async();
                        atomicAdd(p.cnt_a + mt, 1);
                        WS_TRACE(4);
                    }
                }
            } else if (role == R_BH) {
                // partial sums of the layer-1 recurrent half, for the layer-1 input CTA of the same slice
                float4 *dst = reinterpret_cast<float4 *>(p.part + ((((size_t)mt * W_NG + slice) * 2 + cgp) * W_BM + r_in) * 32);
                mbar_wait_wd(&sm.acc_full[buf], use & 1);
                if (etid == 0) WS_TRACE(3);
                tc_fence_after();
                load_acc(taddr, buf, r);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    __stcg(dst + j, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                                __uint_as_float(r[4 * j + 3])));
                if (TS) {  // the unit's stores are issued: the signal thread fences and publishes them; this thread moves on
                    if (etid == 0) WS_TRACE(5);
                    mbar_arrive(&sm.sig_full[(tile - 1) % W_Q]);
                } else {
                    named_bar_sync(1, W_EPI_THREADS);
                    if (etid == 0) {
                        __threadfence();
                        st_release(p.part_ready + mt * W_NG + slice, it + 1);
                        WS_TRACE(4);
                    }
                }
            } else if (role == R_BI) {
                // every load that does not depend on another is issued up front (at full occupancy the epilogue pipeline, one
                // unit at a time, is what bounds a CTA: dependent L2 round trips are its cost): control row + cell state go out
                // together with the acquire of the recurrent partner's flag, the partial sums right after it
                float *cst = p.c1 + (size_t)row * kH + nb / 4;
                const WCtl *crow = p.ctl + (size_t)(it & 1) * p.Mpad + row;  // control rows of step `it` (published by layer-0 slice 0)
                const int4 ctl_a = __ldcg(reinterpret_cast<const int4 *>(crow)), ctl_b = __ldcg(reinterpret_cast<const int4 *>(crow) + 1);
                // every layer-0 epilogue of this step has consumed the argmax keys of step it-1 (that is what released this unit):
                // clear them for the vocabulary phase of step it+1, which reuses the buffer
                if (slice == 0 && cgp == 0) __stcg(p.amax + (size_t)((it + 1) & 1) * p.Mpad + row, 0ull);
                float4 ad[8], cold4[2], pr[8];
                WS_TRACE_EPI(1);
                cold4[0] = __ldcg(reinterpret_cast<const float4 *>(cst));
                cold4[1] = __ldcg(reinterpret_cast<const float4 *>(cst) + 1);
                spin_ge(p.part_ready + mt * W_NG + slice, it + 1);  // per thread: acquire orders the partial-sum loads below
                WS_TRACE_EPI(2);
                {
                    const float4 *src = reinterpret_cast<const float4 *>(p.part + ((((size_t)mt * W_NG + slice) * 2 + cgp) * W_BM + r_in) * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) pr[j] = __ldcg(src + j);
                    const float4 *addp = reinterpret_cast<const float4 *>(p.b1p + nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) ad[j] = __ldg(addp + j);
                }
                WCtl c;
                c.t = ctl_a.x; c.sym = ctl_a.y; c.total = ctl_a.z; c.last = ctl_a.w; c.active = ctl_b.x; c.nsteps = ctl_b.y; c.failed = ctl_b.z; c.pad = ctl_b.w;
                mbar_wait_wd(&sm.acc_full[buf], use & 1);
                if (etid == 0) WS_TRACE(3);
                tc_fence_after();
                load_acc_add(taddr, buf, reinterpret_cast<float *>(pr));  // (accumulator + recurrent partial sums) ...
                WS_TRACE_EPI(6);
                if (c.active) {
                    float cold[8] = {cold4[0].x, cold4[0].y, cold4[0].z, cold4[0].w, cold4[1].x, cold4[1].y, cold4[1].z, cold4[1].w};
                    float hnew[8];
                    __align__(16) __nv_bfloat16 vh[8], vl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {  // ... + bias
                        const float gi = fsig(pr[j].x + ad[j].x), gf = fsig(pr[j].y + ad[j].y);
                        const float gg = ftanh(pr[j].z + ad[j].z), go = fsig(pr[j].w + ad[j].w);
                        const float cn = gf * cold[j] + gi * gg;
                        cold[j] = cn;
                        hnew[j] = go * ftanh(cn);
                        split_bf16(hnew[j], vh[j], vl[j]);
                    }
                    float *hf = p.h1f + (size_t)row * kH + nb / 4;
                    const size_t ob = ((size_t)(par ^ 1) * p.Mpad + row) * kH + nb / 4;
                    reinterpret_cast<float4 *>(cst)[0] = make_float4(cold[0], cold[1], cold[2], cold[3]);
                    reinterpret_cast<float4 *>(cst)[1] = make_float4(cold[4], cold[5], cold[6], cold[7]);
                    reinterpret_cast<float4 *>(hf)[0] = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
                    reinterpret_cast<float4 *>(hf)[1] = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
                    *reinterpret_cast<uint4 *>(p.h1b_hi + ob) = *reinterpret_cast<uint4 *>(vh);
                    *reinterpret_cast<uint4 *>(p.h1b_lo + ob) = *reinterpret_cast<uint4 *>(vl);
                }
                if (TS) {  // the unit's stores are issued: the signal thread fences and publishes them; this thread moves on
                    if (etid == 0) WS_TRACE(5);
                    mbar_arrive(&sm.sig_full[(tile - 1) % W_Q]);
                } else {
                        named_bar_sync(1, W_EPI_THREADS);
                    if (etid == 0) {  // cumulative release of every epilogue thread's stores (ordered by the barrier); the readers use TMA.
                        // (One release per thread instead — no barrier — was measured 3x slower: 10 k atomics per M-tile step on one word.)
                        __threadfence();
                        fence_proxy_async();
                        atomicAdd(p.cnt_b + mt, 1);
                        WS_TRACE(4);
                    }
                }
            } else if (role == R_C) {
                const int4 ri = __ldg(p.rowinfo + row);
                const WCtl c = load_ctl(p.ctl + (size_t)(it & 1) * p.Mpad + row);
                float4 ev[8] = {};
                if (c.active) {
                    const float4 *ep = reinterpret_cast<const float4 *>(p.E + ((size_t)ri.z + c.t) * kH + nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) ev[j] = __ldg(ep + j);
                    // E is streamed from HBM exactly once per (stream, frame): pull the next frame's 128 bytes into L2 now so
                    // that the load above finds them there when the stream advances (this load sits on the step chain)
                    if (c.t + 1 < ri.y) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep + kH / 4));
                }
                mbar_wait_wd(&sm.acc_full[buf], use & 1);
                if (etid == 0) WS_TRACE(3);
                tc_fence_after();
                load_acc_add(taddr, buf, reinterpret_cast<float *>(ev));
                if (c.active) {
                    __nv_bfloat16 *bh = p.zb_hi + (size_t)row * kH + nb, *bl = p.zb_lo + (size_t)row * kH + nb;
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        __align__(16) __nv_bfloat16 vh[8], vl[8];
                        const float e8[8] = {ev[2 * j8].x, ev[2 * j8].y, ev[2 * j8].z, ev[2 * j8].w,
                                             ev[2 * j8 + 1].x, ev[2 * j8 + 1].y, ev[2 * j8 + 1].z, ev[2 * j8 + 1].w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v = e8[j];
                            split_bf16(p.relu ? fmaxf(v, 0.f) : ftanh(v), vh[j], vl[j]);
                        }
                        reinterpret_cast<uint4 *>(bh)[j8] = *reinterpret_cast<uint4 *>(vh);
                        reinterpret_cast<uint4 *>(bl)[j8] = *reinterpret_cast<uint4 *>(vl);
                    }
                }
                if (TS) {  // the unit's stores are issued: the signal thread fences and publishes them; this thread moves on
                    if (etid == 0) WS_TRACE(5);
                    mbar_arrive(&sm.sig_full[(tile - 1) % W_Q]);
                } else {
                    named_bar_sync(1, W_EPI_THREADS);
                    if (etid == 0) {  // cumulative release of every epilogue thread's stores (ordered by the barrier); the readers use TMA.
                        // (One release per thread instead — no barrier — was measured 3x slower: 10 k atomics per M-tile step on one word.)
                        __threadfence();
                        fence_proxy_async();
                        atomicAdd(p.cnt_c + mt, 1);
                        WS_TRACE(4);
                    }
                }
            } else {  // R_D: vocabulary slice -> first-max argmax partial (zero_copy.rs:190-232 tie rule) -> control update
                const WCtl c = load_ctl(p.ctl + (size_t)(it & 1) * p.Mpad + row);
                float4 bo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)  // the padding slice of the cluster variant lies beyond the padded bias vector
                    bo[j] = nb + 32 <= W_ND * W_SL ? __ldg(reinterpret_cast<const float4 *>(p.boutp + nb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                mbar_wait_wd(&sm.acc_full[buf], use & 1);
                if (etid == 0) WS_TRACE(3);
                tc_fence_after();
                load_acc_add(taddr, buf, reinterpret_cast<float *>(bo));
                if (c.active) {
                    float best_v = -INFINITY;
                    int best_i = 0x7fffffff;
                    const float *bof = reinterpret_cast<const float *>(bo);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nb + j;
                        if (n < kV) {
                            // zero_copy.rs:190-232 seeds (max, idx) with (logits[0], 0) and replaces on strict '>': a NaN can
                            // only be returned from index 0 (nothing compares greater than it), a NaN anywhere else never wins
                            const float v = bof[j];
                            if (n == 0 || v > best_v || (best_i == 0x7fffffff && v == v)) { best_v = v; best_i = n; }
                        }
                    }
                    if (best_i != 0x7fffffff) {
                        // order-preserving map of the float, then ~column: the maximum key is the largest logit and, among equal
                        // logits, the smallest column — the strict-'>' first-max rule of zero_copy.rs:190-232.  -0.0 == +0.0
                        // there, so both map to one key; a NaN seed (column 0 only) maps above every number.
                        const unsigned fb = best_v == 0.f ? 0u : __float_as_uint(best_v);
                        const unsigned ord = best_v != best_v ? 0xFFFFFFFFu : ((fb & 0x80000000u) ? ~fb : (fb | 0x80000000u));
                        atomicMax(p.amax + (size_t)(it & 1) * p.Mpad + row, ((unsigned long long)ord << 32) | (0xFFFFFFFFu - (unsigned)best_i));
                    }
                }
                if (TS) {  // the unit's stores are issued: the signal thread fences and publishes them; this thread moves on
                    if (etid == 0) WS_TRACE(5);
                    mbar_arrive(&sm.sig_full[(tile - 1) % W_Q]);
                } else {
                    named_bar_sync(1, W_EPI_THREADS);
                    if (etid == 0) {  // the layer-0 epilogues of the next step read the merged argmax and apply the control flow
                        __threadfence();
                        atomicAdd(p.cnt_d + mt, 1);
                        WS_TRACE(4);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 1) tmem_dealloc(sm.tmem_slot, 512);
}

size_t ws_align(size_t x) { return (x + 1023) & ~(size_t)1023; }

}  // namespace

bool decoder_ws_supported(const Ctx *c) { return c->sm_count >= W_CTAS; }

cudaError_t decoder_ws_prepare(Ctx *c) {
    TcWeights *w = c->dec->tc;
    cudaError_t e;
    if ((e = make_tmap_bf16(&w->s_whh0_hi, w->whh0_hi, kG, kH, kH, W_SL)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->s_whh0_lo, w->whh0_lo, kG, kH, kH, W_SL)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->s_w1_hi, w->w1_hi, kG, 2 * kH, 2 * kH, W_SL)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->s_w1_lo, w->w1_lo, kG, 2 * kH, 2 * kH, W_SL)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->s_wp_hi, w->wp_hi, kH, kH, kH, W_SL)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->s_wp_lo, w->wp_lo, kH, kH, kH, W_SL)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->s_wo_hi, w->wo_hi, W_ND * W_SL, kH, kH, W_SL)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&w->s_wo_lo, w->wo_lo, W_ND * W_SL, kH, kH, W_SL)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(greedy_ws_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM)) != cudaSuccess) return e;
    w->ws_cluster = false;
    w->ws_ready = true;
    return cudaSuccess;
}

// E [sum of encoded lengths][640] fp32 (hoisted encoder projection, valid frames packed; stream b starts at row eoff[b]) is
// produced by the caller (decoder_tc.cu).
cudaError_t launch_greedy_ws(Ctx *c, const float *E, int B, int T, const int32_t *lens_dev, const int *perm_dev,
                             const int *eoff_dev, const int32_t *slots_dev, float *s1_dev, float *s2_dev, int32_t *tokens_dev, int32_t *ntok_dev,
                             int32_t *nsteps_dev, char *work, size_t *work_bytes) {
    DecoderPriv *d = c->dec;
    TcWeights *w = d->tc;
    const int MT = (B + W_BM - 1) / W_BM, Mpad = MT * W_BM;
    const size_t MH = (size_t)Mpad * kH;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += ws_align(bytes); return o; };
    const size_t oh0h = take(2 * 2 * MH), oh0l = take(2 * 2 * MH), oh1h = take(2 * 2 * MH), oh1l = take(2 * 2 * MH);
    const size_t ozh = take(2 * MH), ozl = take(2 * MH);
    const size_t oact_end = off;
    const size_t oh0f = take(4 * MH), oh1f = take(4 * MH), oc0 = take(4 * MH), oc1 = take(4 * MH);
    const size_t opart = take(sizeof(float) * (size_t)MT * W_NG * 2 * W_BM * 32);
    const size_t oamax = take(sizeof(unsigned long long) * 2 * (size_t)Mpad);
    const size_t octl = take(sizeof(WCtl) * 2 * (size_t)Mpad);  // double-buffered by step parity
    const size_t ori = take(sizeof(int4) * (size_t)Mpad);
    const size_t n_cnt = 7 * (size_t)MT + (size_t)MT * W_NG + 4;
    const size_t ocnt = take(sizeof(int) * n_cnt);
    const size_t otrace = take(sizeof(long long) * W_TRACE_ITS * (32 + 64));
    if (!work) {  // size query
        *work_bytes = off;
        return cudaSuccess;
    }
    if (MT > W_MAX_MT || !w || !w->ws_ready) return cudaErrorInvalidValue;
    cudaError_t e;
    if ((e = cudaMemsetAsync(work + ocnt, 0, sizeof(int) * n_cnt, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(work + oamax, 0, sizeof(unsigned long long) * 2 * (size_t)Mpad, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(work + oh0h, 0, oact_end - oh0h, c->stream)) != cudaSuccess) return e;  // padding rows feed the MMA too

    WsParams p;
    std::memset(&p, 0, sizeof(p));
    p.h0b_hi = reinterpret_cast<__nv_bfloat16 *>(work + oh0h); p.h0b_lo = reinterpret_cast<__nv_bfloat16 *>(work + oh0l);
    p.h1b_hi = reinterpret_cast<__nv_bfloat16 *>(work + oh1h); p.h1b_lo = reinterpret_cast<__nv_bfloat16 *>(work + oh1l);
    p.zb_hi = reinterpret_cast<__nv_bfloat16 *>(work + ozh); p.zb_lo = reinterpret_cast<__nv_bfloat16 *>(work + ozl);
    const bool cluster = false;
    // one M-tile with few streams (the reference's B = 1 request, small micro-batches): a unit loads and multiplies only
    // the first 32 / 64 rows of the tile — the activation ingest (327 KB per unit at 128 rows) is what a chain phase waits for
    p.nrows = (MT == 1 && !getenv("AMIRA_WS_FULLROWS")) ? (B <= 32 ? 32 : B <= 64 ? 64 : W_BM) : W_BM;
    const uint32_t box_rows = cluster ? W_BM / 2 : (uint32_t)p.nrows;  // the cluster variant loads half a tile per CTA and multicasts it
    if ((e = make_tmap_bf16(&p.h0_hi, p.h0b_hi, 2 * (uint64_t)Mpad, kH, kH, box_rows)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.h0_lo, p.h0b_lo, 2 * (uint64_t)Mpad, kH, kH, box_rows)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.h1_hi, p.h1b_hi, 2 * (uint64_t)Mpad, kH, kH, box_rows)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.h1_lo, p.h1b_lo, 2 * (uint64_t)Mpad, kH, kH, box_rows)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.z_hi, p.zb_hi, (uint64_t)Mpad, kH, kH, box_rows)) != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&p.z_lo, p.zb_lo, (uint64_t)Mpad, kH, kH, box_rows)) != cudaSuccess) return e;
    p.whh0_hi = w->s_whh0_hi; p.whh0_lo = w->s_whh0_lo; p.w1_hi = w->s_w1_hi; p.w1_lo = w->s_w1_lo;
    p.wp_hi = w->s_wp_hi; p.wp_lo = w->s_wp_lo; p.wo_hi = w->s_wo_hi; p.wo_lo = w->s_wo_lo;
    p.g0p = d->g0p; p.b1p = d->b1p; p.boutp = d->boutp; p.E = E;
    p.g_whh0_hi = w->whh0_hi; p.g_whh0_lo = w->whh0_lo; p.g_w1_hi = w->w1_hi; p.g_w1_lo = w->w1_lo;
    p.g_wp_hi = w->wp_hi; p.g_wp_lo = w->wp_lo; p.g_wo_hi = w->wo_hi; p.g_wo_lo = w->wo_lo;
    p.B = B; p.Mpad = Mpad; p.MT = MT; p.T = T > 0 ? T : 1;
    p.lens = lens_dev; p.slots = slots_dev; p.perm = perm_dev; p.eoff = eoff_dev;
    p.h0f = reinterpret_cast<float *>(work + oh0f); p.h1f = reinterpret_cast<float *>(work + oh1f);
    p.c0 = reinterpret_cast<float *>(work + oc0); p.c1 = reinterpret_cast<float *>(work + oc1);
    p.part = reinterpret_cast<float *>(work + opart);
    p.amax = reinterpret_cast<unsigned long long *>(work + oamax);
    p.ctl = reinterpret_cast<WCtl *>(work + octl);
    p.rowinfo = reinterpret_cast<int4 *>(work + ori);
    int *cnt = reinterpret_cast<int *>(work + ocnt);
    p.tile_active = cnt; p.cnt_d = cnt + MT; p.cnt_a = cnt + 2 * MT; p.cnt_b = cnt + 3 * MT; p.cnt_c = cnt + 4 * MT;
    p.dead_at = cnt + 6 * MT; p.part_ready = cnt + 7 * MT; p.fail_count = cnt + 7 * MT + MT * W_NG;
    if (slots_dev) { p.s1 = c->slot_s1; p.s2 = c->slot_s2; } else { p.s1 = s1_dev; p.s2 = s2_dev; }
    p.tokens = tokens_dev; p.ntok = ntok_dev; p.nsteps = nsteps_dev;
    p.max_sym = c->cfg.max_symbols_per_step; p.max_total = c->cfg.max_total_tokens; p.blank = c->cfg.blank_id;
    p.relu = c->cfg.joint_activation;
    d->fail_count_dev = p.fail_count;
    p.norot = getenv("AMIRA_WS_NOROT") ? 1 : 0;
    p.variant = getenv("AMIRA_WS_VARIANT") ? atoi(getenv("AMIRA_WS_VARIANT")) : 3;
    if (getenv("AMIRA_WS_TRACE")) {
        p.trace = reinterpret_cast<long long *>(work + otrace);
        cudaMemsetAsync(p.trace, 0, sizeof(long long) * W_TRACE_ITS * (32 + 64), c->stream);
        p.trace_role = getenv("AMIRA_WS_TRACE_ROLE") ? atoi(getenv("AMIRA_WS_TRACE_ROLE")) : R_BI;
        d->ws_trace_dev = p.trace;
    }

    void *params[] = {&p};
    ProfScope prof(c, PK_GREEDY);
    e = cudaLaunchCooperativeKernel((const void *)greedy_ws_kernel<1, true>, dim3(W_CTAS), dim3(W_THREADS), params, W_SMEM, c->stream);
    c->launches++;
    return e;
}

}  // namespace amira

// ---- diagnostics: globaltimer stamps of the last weight-stationary launch (set AMIRA_WS_TRACE=1 before the call) ----
extern "C" int32_t amira_debug_ws_trace(amira_ctx *ctx, int64_t *out, int32_t n_its) {
    using namespace amira;
    if (!ctx || !out || n_its <= 0 || (n_its > W_TRACE_ITS && n_its != 3 * W_TRACE_ITS)) return AMIRA_ERR_INVALID_VALUE;
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    std::lock_guard<std::mutex> lock(c->mu);
    if (!c->dec || !c->dec->ws_trace_dev) return AMIRA_ERR_NOT_READY;
    cudaSetDevice(c->device);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return AMIRA_ERR_UNKNOWN;
    return cudaMemcpy(out, c->dec->ws_trace_dev, sizeof(int64_t) * 32 * (size_t)n_its, cudaMemcpyDeviceToHost) == cudaSuccess
               ? AMIRA_OK : AMIRA_ERR_UNKNOWN;
}
