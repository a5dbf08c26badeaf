// decoder.cu — the `decoder_joint` stage on B200 (sm_100a): the RNN-T greedy loop as ONE persistent cooperative
// kernel, plus the Triton-contract op.
//
// Replaces (citations relative to the reference root):
//   greedy_decode / greedy_decode_zero_copy            src/asr/decoder_optimized.rs:24-200
//   the per-step decode closure + DecoderJointModel    src/asr/pipeline.rs:313-356, src/triton/model.rs:581-722
//   TensorView::extract_frame_into, argmax_zero_copy   src/asr/zero_copy.rs:49-69,190-232
// The reference makes one gRPC round trip per decode step and B = 1; here B streams advance in lock step
// inside one kernel and nothing returns to the host until every stream is finished.
//
// Algebra (same function as the reference's model, re-associated):
//   layer-0 gates  = G0[token] + W_hh0 h0           G0 = emb W_ih0^T + b_ih0 + b_hh0  (table, built at load time)
//   layer-1 gates  = b1 + [W_ih1 | W_hh1] [h0'; h1]
//   joint hidden   = act(E[b][t] + W_pred h1')      E = enc W_enc^T + b_enc + b_pred  (hoisted: once per frame)
//   logits         = W_out z + b_out -> first-max argmax over all 1030 outputs (zero_copy.rs:190-232)
// Gate columns are permuted to unit-major (u*4 + gate) so the thread that finishes a 4-column micro-tile owns
// one hidden unit and applies the cell update in registers.
//
// Per iteration the kernel runs four grid-synchronised phases (L0, L1, joint-hidden, vocab+argmax); the
// per-stream control flow of decoder_optimized.rs:88-188 (blank advance, <= max_symbols step calls per frame,
// max_total_tokens, unconditional state carry) is applied on-device at the top of the next iteration.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>

#include "common.h"

namespace cg = cooperative_groups;

namespace amira {

namespace {

constexpr int TM = 64, TN = 64, TK = 16, LDS_ = TM + 4;
constexpr int DEC_THREADS = 256;
constexpr int NT_GATES = kG / TN;        // 40
constexpr int NT_PRED = kH / TN;         // 10
constexpr int NT_OUT = (kV + TN - 1) / TN;  // 17
constexpr int V_PAD = NT_OUT * TN;       // 1088

struct Ctl {
    int t, sym, total, last, active, par, nsteps, failed;
};

struct DecWeights {
    const float *g0p, *whh0p, *w1p, *b1p, *wpred, *woutp, *boutp;
};

struct DecArgs {
    DecWeights w;
    const float *E;       // [B][T][640]
    int B, T;
    const int *lens;      // [B]
    const int *slots;     // nullable
    float *h0, *h1;       // [2][B][640] ping-pong by row parity
    float *c0, *c1;       // [2][B][640] ping-pong by row parity, like h0 / h1 (a step that is not carried leaves no trace)
    float *z;             // [B][640]
    float *pval;          // [B][NT_OUT]
    int *pidx;            // [B][NT_OUT]
    Ctl *ctl;             // [2][B]
    int *act_count;       // [0..1] active rows per parity, [2] failed streams
    float *s1, *s2;       // in/out states (nullable); batch layout [2][B][640] or slot layout [slot][2][640]
    int *tokens, *ntok, *nsteps;
    int max_sym, max_total, blank, relu;
    int rule;             // AMIRA_RULE_* bits (0 = the reference's literal loop)
    float *dur;           // [B][8] duration logits (outputs blank+1 .. 1029) of the last step, TDT reading only
};

struct TileSmem {
    float As[2][TK][LDS_];
    float Ws[2][TK][LDS_];
    const float *rowA[2][TM];
    Ctl ctl[TM];
    int any;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// C[64x64] = A[64 x (nseg*640)] * W[64 x (nseg*640)]^T ; A rows through s.rowA[seg][r] (640 contiguous floats each),
// W rows contiguous with row stride ldw.  256 threads, 4x4 micro-tile per thread, register-prefetched smem
// double buffer.
__device__ __forceinline__ void gemm_tile(TileSmem &s, int nseg, const float *__restrict__ W, int ldw, float (&acc)[4][4]) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row lr, k offset lk
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int chunks_per_seg = kH / TK;  // 40
    const int nchunk = nseg * chunks_per_seg;
    float4 ra, rw;
    auto fetch = [&](int ch) {
        const int seg = ch / chunks_per_seg, k0 = (ch % chunks_per_seg) * TK;
        ra = *reinterpret_cast<const float4 *>(s.rowA[seg][lr] + k0 + lk);
        rw = __ldg(reinterpret_cast<const float4 *>(W + (size_t)lr * ldw + seg * kH + k0 + lk));
    };
    auto stash = [&](int buf) {
        s.As[buf][lk + 0][lr] = ra.x; s.As[buf][lk + 1][lr] = ra.y; s.As[buf][lk + 2][lr] = ra.z; s.As[buf][lk + 3][lr] = ra.w;
        s.Ws[buf][lk + 0][lr] = rw.x; s.Ws[buf][lk + 1][lr] = rw.y; s.Ws[buf][lk + 2][lr] = rw.z; s.Ws[buf][lk + 3][lr] = rw.w;
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int ch = 0; ch < nchunk; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunk) fetch(ch + 1);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&s.As[buf][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&s.Ws[buf][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (ch + 1 < nchunk) stash(buf ^ 1);
        __syncthreads();
    }
}

// LSTM cell update for one (row, unit): gates pre-activation in a[4] (i, f, g, o)
__device__ __forceinline__ void lstm_cell(const float (&a)[4], const float *c_old, float *c_new, float *h_ptr) {
    const float ig = sigmoidf_(a[0]), fg = sigmoidf_(a[1]), gg = tanhf(a[2]), og = sigmoidf_(a[3]);
    const float cn = fg * (*c_old) + ig * gg;
    *c_new = cn;
    *h_ptr = og * tanhf(cn);
}

__device__ __forceinline__ size_t state_off(const DecArgs &a, int layer, int b) {
    return a.slots ? ((size_t)a.slots[b] * 2 + layer) * kH : ((size_t)layer * a.B + b) * kH;
}

__global__ void __launch_bounds__(DEC_THREADS, 2) greedy_persistent_kernel(DecArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ TileSmem s;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int B = a.B;
    const int MT = (B + TM - 1) / TM;
    const size_t BH = (size_t)B * kH;

    // ---- prologue: load initial state, reset control ----
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < BH; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / kH), j = (int)(i % kH);
        a.h0[i] = a.s1 ? a.s1[state_off(a, 0, b) + j] : 0.f;
        a.h1[i] = a.s1 ? a.s1[state_off(a, 1, b) + j] : 0.f;
        a.c0[i] = a.s2 ? a.s2[state_off(a, 0, b) + j] : 0.f;
        a.c1[i] = a.s2 ? a.s2[state_off(a, 1, b) + j] : 0.f;
    }
    for (int b = blockIdx.x * blockDim.x + tid; b < B; b += gridDim.x * blockDim.x) {
        Ctl c;
        c.t = 0; c.sym = 0; c.total = 0; c.last = a.blank; c.active = a.lens[b] > 0 ? 1 : 0; c.par = 0; c.nsteps = 0; c.failed = 0;
        a.ctl[B + b] = c;  // parity 1 = "previous" of iteration 0
    }
    grid.sync();

    int p = 0;
    for (int it = 0;; ++it) {
        p = it & 1;
        const Ctl *prev = a.ctl + (size_t)(1 - p) * B;
        Ctl *cur = a.ctl + (size_t)p * B;

        // ================= phase A: control update + layer 0 =================
        for (int tile = blockIdx.x; tile < MT * NT_GATES; tile += gridDim.x) {
            const int mt = tile / NT_GATES, nt = tile % NT_GATES;
            __syncthreads();
            if (tid == 0) s.any = 0;
            __syncthreads();
            if (tid < TM) {
                const int row = mt * TM + tid;
                Ctl c{};
                if (row < B) {
                    c = prev[row];
                    if (it > 0 && c.active) {  // a step call was made for this row in the previous iteration
                        float bv = a.pval[(size_t)row * NT_OUT];
                        int bi = a.pidx[(size_t)row * NT_OUT];
                        for (int q = 1; q < NT_OUT; ++q) {
                            const float v = a.pval[(size_t)row * NT_OUT + q];
                            if (v > bv) { bv = v; bi = a.pidx[(size_t)row * NT_OUT + q]; }
                        }
                        // One rule set for the reference's loop and the two non-reference variants of SURVEY 8(f4) (a.rule):
                        // literal (rule 0): state carried unconditionally, blank advances one frame; AMIRA_RULE_STATE_ON_NONBLANK: a
                        // blank leaves the prediction net where it was (the step's results stay in the other parity and are
                        // overwritten); AMIRA_RULE_TDT_DURATIONS: the token is the first max over outputs [0, blank], the frame
                        // advance the first max over the duration outputs (blank with duration 0 advances one frame).
                        const bool is_blank = bi == a.blank;
                        c.nsteps += 1;
                        if (!(is_blank && (a.rule & AMIRA_RULE_STATE_ON_NONBLANK))) c.par ^= 1;  // decoder_optimized.rs:154 when literal
                        c.sym += 1;      // :133
                        int skip = -1;
                        if (a.rule & AMIRA_RULE_TDT_DURATIONS) {
                            const float *dv = a.dur + (size_t)row * 8;
                            float best = dv[0];
                            skip = 0;
                            for (int q = 1; q < kV - a.blank - 1 && q < 8; ++q)
                                if (dv[q] > best) { best = dv[q]; skip = q; }
                            if (is_blank && skip == 0) skip = 1;
                        }
                        const int len = a.lens[row];
                        if (!is_blank) {
                            if (nt == 0) a.tokens[(size_t)row * a.max_total + c.total] = bi;   // :176
                            c.total += 1;
                            c.last = bi;
                            if (c.total >= a.max_total) c.active = 0;                           // :179-188
                        }
                        if (c.active) {
                            int adv = skip >= 0 ? skip : (is_blank ? 1 : 0);                    // :171-173
                            if (adv == 0 && c.sym >= a.max_sym) adv = 1;                        // :133-137
                            if (adv) {
                                c.t += adv; c.sym = 0;
                                if (c.t >= len) c.active = 0;
                            }
                            // an id outside the embedding table fails the next step call ("Decode step failed", :148-152)
                            if (c.active && bi >= kEmbRows) { c.active = 0; c.failed = 1; }
                        }
                    }
                    if (nt == 0) cur[row] = c;
                }
                s.ctl[tid] = c;
                if (c.active) s.any = 1;
                const int rr = (row < B) ? row : 0;
                s.rowA[0][tid] = a.h0 + (size_t)(c.active ? c.par : 0) * BH + (size_t)rr * kH;
            }
            if (nt == 0) {
                // one counter update per M-tile
                __syncthreads();
                if (tid == 0) {
                    int n = 0;
                    for (int r = 0; r < TM; ++r) n += s.ctl[r].active;
                    if (n) atomicAdd(&a.act_count[p], n);
                }
            }
            __syncthreads();
            if (!s.any) continue;
            float acc[4][4];
            gemm_tile(s, 1, a.w.whh0p + (size_t)nt * TN * kH, kH, acc);
            const int u = nt * (TN / 4) + tx;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty * 4 + i, row = mt * TM + r;
                const Ctl c = s.ctl[r];
                if (row < B && c.active) {
                    const float4 g = __ldg(reinterpret_cast<const float4 *>(a.w.g0p + (size_t)c.last * kG + nt * TN + tx * 4));
                    const float pre[4] = {acc[i][0] + g.x, acc[i][1] + g.y, acc[i][2] + g.z, acc[i][3] + g.w};
                    lstm_cell(pre, a.c0 + (size_t)c.par * BH + (size_t)row * kH + u, a.c0 + (size_t)(c.par ^ 1) * BH + (size_t)row * kH + u,
                              a.h0 + (size_t)(c.par ^ 1) * BH + (size_t)row * kH + u);
                }
            }
        }
        grid.sync();
        if (*((volatile int *)&a.act_count[p]) == 0) break;
        if (blockIdx.x == 0 && tid == 0) a.act_count[1 - p] = 0;

        // ================= phase B: layer 1 =================
        for (int tile = blockIdx.x; tile < MT * NT_GATES; tile += gridDim.x) {
            const int mt = tile / NT_GATES, nt = tile % NT_GATES;
            __syncthreads();
            if (tid == 0) s.any = 0;
            __syncthreads();
            if (tid < TM) {
                const int row = mt * TM + tid;
                Ctl c{};
                if (row < B) c = cur[row];
                s.ctl[tid] = c;
                if (c.active) s.any = 1;
                const int rr = (row < B) ? row : 0;
                s.rowA[0][tid] = a.h0 + (size_t)(c.active ? (c.par ^ 1) : 0) * BH + (size_t)rr * kH;
                s.rowA[1][tid] = a.h1 + (size_t)(c.active ? c.par : 0) * BH + (size_t)rr * kH;
            }
            __syncthreads();
            if (!s.any) continue;
            float acc[4][4];
            gemm_tile(s, 2, a.w.w1p + (size_t)nt * TN * 2 * kH, 2 * kH, acc);
            const int u = nt * (TN / 4) + tx;
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.w.b1p + nt * TN + tx * 4));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty * 4 + i, row = mt * TM + r;
                const Ctl c = s.ctl[r];
                if (row < B && c.active) {
                    const float pre[4] = {acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w};
                    lstm_cell(pre, a.c1 + (size_t)c.par * BH + (size_t)row * kH + u, a.c1 + (size_t)(c.par ^ 1) * BH + (size_t)row * kH + u,
                              a.h1 + (size_t)(c.par ^ 1) * BH + (size_t)row * kH + u);
                }
            }
        }
        grid.sync();

        // ================= phase C: joint hidden z = act(E[b][t] + W_pred h1') =================
        for (int tile = blockIdx.x; tile < MT * NT_PRED; tile += gridDim.x) {
            const int mt = tile / NT_PRED, nt = tile % NT_PRED;
            __syncthreads();
            if (tid == 0) s.any = 0;
            __syncthreads();
            if (tid < TM) {
                const int row = mt * TM + tid;
                Ctl c{};
                if (row < B) c = cur[row];
                s.ctl[tid] = c;
                if (c.active) s.any = 1;
                const int rr = (row < B) ? row : 0;
                s.rowA[0][tid] = a.h1 + (size_t)(c.active ? (c.par ^ 1) : 0) * BH + (size_t)rr * kH;
            }
            __syncthreads();
            if (!s.any) continue;
            float acc[4][4];
            gemm_tile(s, 1, a.w.wpred + (size_t)nt * TN * kH, kH, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty * 4 + i, row = mt * TM + r;
                const Ctl c = s.ctl[r];
                if (row < B && c.active) {
                    const float4 e = *reinterpret_cast<const float4 *>(a.E + ((size_t)row * a.T + c.t) * kH + nt * TN + tx * 4);
                    float4 zz = make_float4(acc[i][0] + e.x, acc[i][1] + e.y, acc[i][2] + e.z, acc[i][3] + e.w);
                    if (a.relu) {
                        zz.x = fmaxf(zz.x, 0.f); zz.y = fmaxf(zz.y, 0.f); zz.z = fmaxf(zz.z, 0.f); zz.w = fmaxf(zz.w, 0.f);
                    } else {
                        zz.x = tanhf(zz.x); zz.y = tanhf(zz.y); zz.z = tanhf(zz.z); zz.w = tanhf(zz.w);
                    }
                    *reinterpret_cast<float4 *>(a.z + (size_t)row * kH + nt * TN + tx * 4) = zz;
                }
            }
        }
        grid.sync();

        // ================= phase D: logits tile + partial first-max argmax =================
        for (int tile = blockIdx.x; tile < MT * NT_OUT; tile += gridDim.x) {
            const int mt = tile / NT_OUT, nt = tile % NT_OUT;
            __syncthreads();
            if (tid == 0) s.any = 0;
            __syncthreads();
            if (tid < TM) {
                const int row = mt * TM + tid;
                Ctl c{};
                if (row < B) c = cur[row];
                s.ctl[tid] = c;
                if (c.active) s.any = 1;
                const int rr = (row < B) ? row : 0;
                s.rowA[0][tid] = a.z + (size_t)rr * kH;
            }
            __syncthreads();
            if (!s.any) continue;
            float acc[4][4];
            gemm_tile(s, 1, a.w.woutp + (size_t)nt * TN * kH, kH, acc);
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.w.boutp + nt * TN + tx * 4));
            const float bbv[4] = {bb.x, bb.y, bb.z, bb.w};
            const int n_tok = (a.rule & AMIRA_RULE_TDT_DURATIONS) ? a.blank + 1 : kV;  // outputs that compete for the token
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float bv = -INFINITY;
                int bi = 0x7fffffff;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = nt * TN + tx * 4 + j;
                    const float v = acc[i][j] + bbv[j];
                    // zero_copy.rs:190-232: seed (logits[0], 0), replace on strict '>': NaN is returned only from index 0
                    if (n < n_tok && (n == 0 || v > bv || (bi == 0x7fffffff && v == v))) { bv = v; bi = n; }
                    // TDT reading: the outputs behind blank are duration logits, kept for the control update
                    if (n >= n_tok && n < kV && n - n_tok < 8) {
                        const int r_ = ty * 4 + i, row_ = mt * TM + r_;
                        if (row_ < B && s.ctl[r_].active) a.dur[(size_t)row_ * 8 + (n - n_tok)] = v;
                    }
                }
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
                }
                const int r = ty * 4 + i, row = mt * TM + r;
                if (tx == 0 && row < B && s.ctl[r].active) {
                    a.pval[(size_t)row * NT_OUT + nt] = bv;
                    a.pidx[(size_t)row * NT_OUT + nt] = bi;
                }
            }
        }
        grid.sync();
    }

    // ---- epilogue: results ----
    const Ctl *fin = a.ctl + (size_t)p * B;
    for (int b = blockIdx.x * blockDim.x + tid; b < B; b += gridDim.x * blockDim.x) {
        const Ctl c = fin[b];
        a.ntok[b] = c.failed ? -1 : c.total;
        if (c.failed) atomicAdd(&a.act_count[2], 1);
        if (a.nsteps) a.nsteps[b] = c.nsteps;
    }
    if (a.s1 && a.s2) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < BH; i += (size_t)gridDim.x * blockDim.x) {
            const int b = (int)(i / kH), j = (int)(i % kH);
            const int par = fin[b].par;
            a.s1[state_off(a, 0, b) + j] = a.h0[(size_t)par * BH + i];
            a.s1[state_off(a, 1, b) + j] = a.h1[(size_t)par * BH + i];
            a.s2[state_off(a, 0, b) + j] = a.c0[(size_t)par * BH + i];
            a.s2[state_off(a, 1, b) + j] = a.c1[(size_t)par * BH + i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic 64x64 tile GEMM kernel used at load time (G0 table), for the hoisted encoder projection and by the
// contract op: C[m][n] = sum_k A(m,k) * W[n][k] (+ bias[n]) with A either row-major (lda) or k-major
// ("transposed": A(m,k) = A[k*lda + m], the [1024][T] layout of encoder_outputs, zero_copy.rs:61-62).
// grid = (ceil(N/64), ceil(M/64), batch); per-batch strides for A and C; rows >= m_valid[batch] are skipped.
struct GemmArgs {
    const float *A; size_t a_batch; int lda; int a_kmajor;
    const float *W; int ldw;            // [N][K]
    const float *bias;                  // [N] nullable
    const float *bias2;                 // [N] nullable
    float *C; size_t c_batch; int ldc;
    int M, N, K;
    const int *m_valid;                 // per-batch row limit (nullable => M)
};

__global__ void __launch_bounds__(DEC_THREADS) gemm_nt_kernel(GemmArgs g) {
    __shared__ float As[TK][LDS_];
    __shared__ float Ws[TK][LDS_];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int bz = blockIdx.z, m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int mlim = g.m_valid ? min(g.m_valid[bz], g.M) : g.M;
    if (m0 >= mlim) return;
    const float *A = g.A + (size_t)bz * g.a_batch;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < g.K; k0 += TK) {
        // A tile
        if (g.a_kmajor) {
            for (int e = tid; e < TK * TM; e += DEC_THREADS) {
                const int k = e / TM, m = e % TM;
                As[k][m] = (m0 + m < g.M && k0 + k < g.K) ? A[(size_t)(k0 + k) * g.lda + m0 + m] : 0.f;
            }
        } else {
            for (int e = tid; e < TK * TM; e += DEC_THREADS) {
                const int m = e / TK, k = e % TK;
                As[k][m] = (m0 + m < g.M && k0 + k < g.K) ? A[(size_t)(m0 + m) * g.lda + k0 + k] : 0.f;
            }
        }
        for (int e = tid; e < TK * TN; e += DEC_THREADS) {
            const int n = e / TK, k = e % TK;
            Ws[k][n] = (n0 + n < g.N && k0 + k < g.K) ? __ldg(g.W + (size_t)(n0 + n) * g.ldw + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Ws[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *C = g.C + (size_t)bz * g.c_batch;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= mlim) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < g.N) {
                float v = acc[i][j];
                if (g.bias) v += g.bias[n];
                if (g.bias2) v += g.bias2[n];
                C[(size_t)m * g.ldc + n] = v;
            }
        }
    }
}

cudaError_t run_gemm(Ctx *c, const GemmArgs &g, int batch) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return cudaSuccess;
    dim3 grid((g.N + TN - 1) / TN, (g.M + TM - 1) / TM, batch);
    gemm_nt_kernel<<<grid, DEC_THREADS, 0, c->stream>>>(g);
    c->launches++;
    return cudaGetLastError();
}

// ---- load-time permutations ----
__global__ void permute_gate_rows_kernel(const float *__restrict__ w_a, const float *__restrict__ w_b, int ka, int kb,
                                         float *__restrict__ out) {
    // out[p][0..ka) = w_a[row(p)][:], out[p][ka..ka+kb) = w_b[row(p)][:], p = u*4 + g, row = g*640 + u
    const int p = blockIdx.x, u = p >> 2, g = p & 3, row = g * kH + u;
    for (int k = threadIdx.x; k < ka; k += blockDim.x) out[(size_t)p * (ka + kb) + k] = w_a[(size_t)row * ka + k];
    for (int k = threadIdx.x; k < kb; k += blockDim.x) out[(size_t)p * (ka + kb) + ka + k] = w_b[(size_t)row * kb + k];
}
__global__ void permute_gate_bias_kernel(const float *__restrict__ b_a, const float *__restrict__ b_b, float *__restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < kG) {
        const int row = (p & 3) * kH + (p >> 2);
        out[p] = b_a[row] + b_b[row];
    }
}
__global__ void pad_out_kernel(const float *__restrict__ w_out, const float *__restrict__ b_out, float *__restrict__ wp,
                               float *__restrict__ bp) {
    const int n = blockIdx.x;
    for (int k = threadIdx.x; k < kH; k += blockDim.x) wp[(size_t)n * kH + k] = n < kV ? w_out[(size_t)n * kH + k] : 0.f;
    if (threadIdx.x == 0) bp[n] = n < kV ? b_out[n] : 0.f;
}
__global__ void add_vec_kernel(const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ o, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = x[i] + y[i];
}

// ---- contract op helpers ----
// one LSTM layer step for B rows (all rows with u < tlen[b] advance); gates = addend + A W^T
struct StepArgs {
    const float *W; int ldw; int nseg;
    const float *a0, *a1;        // A segments, row stride 640
    const float *g0p;            // layer 0: gather table (nullable)
    const int *targets; int U, u;
    const int *tlen;             // nullable => U
    const float *bias;           // layer 1 bias (permuted)
    float *c;                    // [B][640] in place
    const float *h_old; float *h_new;  // [B][640]
    int B;
    int *err_flag;
};
__global__ void __launch_bounds__(DEC_THREADS) lstm_step_kernel(StepArgs q) {
    __shared__ TileSmem s;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int nt = blockIdx.x, mt = blockIdx.y;
    if (tid < TM) {
        const int row = mt * TM + tid, rr = row < q.B ? row : 0;
        s.rowA[0][tid] = q.a0 + (size_t)rr * kH;
        s.rowA[1][tid] = (q.nseg > 1 ? q.a1 : q.a0) + (size_t)rr * kH;
    }
    __syncthreads();
    float acc[4][4];
    gemm_tile(s, q.nseg, q.W + (size_t)nt * TN * q.ldw, q.ldw, acc);
    const int un = nt * (TN / 4) + tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = mt * TM + ty * 4 + i;
        if (row >= q.B) continue;
        const int tl = q.tlen ? q.tlen[row] : q.U;
        float *hn = q.h_new + (size_t)row * kH + un;
        if (q.u >= tl) {  // position beyond target_length: state frozen
            *hn = q.h_old[(size_t)row * kH + un];
            continue;
        }
        float add[4];
        if (q.g0p) {
            int tok = q.targets[(size_t)row * q.U + q.u];
            if (tok < 0 || tok >= kEmbRows) {  // ONNX Gather out of range in the reference => failed request
                if (tx == 0 && nt == 0) atomicExch(q.err_flag, 1);
                tok = AMIRA_BLANK_ID;
            }
            const float4 g = __ldg(reinterpret_cast<const float4 *>(q.g0p + (size_t)tok * kG + nt * TN + tx * 4));
            add[0] = g.x; add[1] = g.y; add[2] = g.z; add[3] = g.w;
        } else {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(q.bias + nt * TN + tx * 4));
            add[0] = g.x; add[1] = g.y; add[2] = g.z; add[3] = g.w;
        }
        const float pre[4] = {acc[i][0] + add[0], acc[i][1] + add[1], acc[i][2] + add[2], acc[i][3] + add[3]};
        lstm_cell(pre, q.c + (size_t)row * kH + un, q.c + (size_t)row * kH + un, hn);
    }
}

// z[((b*U + u)*T + t)][j] = act(E[b][t][j] + P[b][u][j])
__global__ void joint_hidden_kernel(const float *__restrict__ E, const float *__restrict__ P, float *__restrict__ z, int B,
                                    int U, int T, int relu) {
    const size_t n = (size_t)B * U * T * kH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % kH);
        const size_t r = i / kH;
        const int t = (int)(r % T);
        const size_t bu = r / T;
        const int b = (int)(bu / U);
        const float v = E[((size_t)b * T + t) * kH + j] + P[bu * kH + j];
        z[i] = relu ? fmaxf(v, 0.f) : tanhf(v);
    }
}
__global__ void fill_int_kernel(int *p, const int *src, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = src ? src[i] : v;
}
// [2][B][640] <-> per-layer [B][640] copies are plain offsets; nothing to do.

}  // namespace

const int32_t *decoder_fail_count_dev(Ctx *c) { return c->dec ? c->dec->fail_count_dev : nullptr; }

void decoder_release(Ctx *c) {
    if (!c->dec) return;
    c->dec->work.release();
    delete c->dec;
    c->dec = nullptr;
}

SharedDev::~SharedDev() {
    cudaSetDevice(device);
    for (float *p : {g0p, whh0p, w1p, b1p, bjoint, woutp, boutp, w_blob})
        if (p) cudaFree(p);
    if (tables_dev) cudaFree(tables_dev);
    decoder_tc_free(tc);
    cudaGetLastError();
}

void decoder_adopt_shared(Ctx *c) {
    SharedDev *sh = c->shared.get();
    if (!sh || sh->version == 0) return;
    if (!c->dec) c->dec = new DecoderPriv();
    DecoderPriv *d = c->dec;
    d->g0p = sh->g0p; d->whh0p = sh->whh0p; d->w1p = sh->w1p; d->b1p = sh->b1p; d->bjoint = sh->bjoint; d->woutp = sh->woutp; d->boutp = sh->boutp;
    d->tc = sh->tc;
    d->coop_blocks_per_sm = sh->coop_blocks_per_sm;
    c->w_blob = sh->w_blob;
    c->tables_dev = sh->tables_dev;
    c->weights_version = sh->version;
    c->has_weights = true;
}

cudaError_t decoder_prepare_weights(Ctx *c) {
    SharedDev *d = c->shared.get();  // the tables are allocated once and rewritten in place by later loads
    if (!c->dec) c->dec = new DecoderPriv();
    const BlobLayout L = blob_layout();
    const float *w = c->w_blob;
    cudaError_t e;
    auto alloc = [&](float **p, size_t n) -> cudaError_t { return *p ? cudaSuccess : cudaMalloc(p, sizeof(float) * n); };
    if ((e = alloc(&d->g0p, (size_t)kEmbRows * kG)) != cudaSuccess) return e;
    if ((e = alloc(&d->whh0p, (size_t)kG * kH)) != cudaSuccess) return e;
    if ((e = alloc(&d->w1p, (size_t)kG * 2 * kH)) != cudaSuccess) return e;
    if ((e = alloc(&d->b1p, kG)) != cudaSuccess) return e;
    if ((e = alloc(&d->bjoint, kH)) != cudaSuccess) return e;
    if ((e = alloc(&d->woutp, (size_t)V_PAD * kH)) != cudaSuccess) return e;
    if ((e = alloc(&d->boutp, V_PAD)) != cudaSuccess) return e;
    // permuted layer-0 input weights + bias, used once to build the G0 table
    float *wih0p = nullptr, *b0p = nullptr;
    if ((e = cudaMalloc(&wih0p, sizeof(float) * (size_t)kG * kH)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b0p, sizeof(float) * kG)) != cudaSuccess) { cudaFree(wih0p); return e; }
    permute_gate_rows_kernel<<<kG, 128, 0, c->stream>>>(w + L.w_ih[0], nullptr, kH, 0, wih0p);
    permute_gate_rows_kernel<<<kG, 128, 0, c->stream>>>(w + L.w_hh[0], nullptr, kH, 0, d->whh0p);
    permute_gate_rows_kernel<<<kG, 128, 0, c->stream>>>(w + L.w_ih[1], w + L.w_hh[1], kH, kH, d->w1p);
    permute_gate_bias_kernel<<<(kG + 255) / 256, 256, 0, c->stream>>>(w + L.b_ih[0], w + L.b_hh[0], b0p);
    permute_gate_bias_kernel<<<(kG + 255) / 256, 256, 0, c->stream>>>(w + L.b_ih[1], w + L.b_hh[1], d->b1p);
    pad_out_kernel<<<V_PAD, 128, 0, c->stream>>>(w + L.w_out, w + L.b_out, d->woutp, d->boutp);
    add_vec_kernel<<<(kH + 255) / 256, 256, 0, c->stream>>>(w + L.b_enc, w + L.b_pred, d->bjoint, kH);
    c->launches += 7;
    GemmArgs g{};
    g.A = w + L.emb; g.a_batch = 0; g.lda = kH; g.a_kmajor = 0;
    g.W = wih0p; g.ldw = kH; g.bias = b0p; g.bias2 = nullptr;
    g.C = d->g0p; g.c_batch = 0; g.ldc = kG;
    g.M = kEmbRows; g.N = kG; g.K = kH; g.m_valid = nullptr;
    e = run_gemm(c, g, 1);
    cudaError_t e2 = cudaStreamSynchronize(c->stream);
    cudaFree(wih0p);
    cudaFree(b0p);
    if (e != cudaSuccess) return e;
    if (e2 != cudaSuccess) return e2;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, greedy_persistent_kernel, DEC_THREADS, 0);
    if (e != cudaSuccess) return e;
    d->coop_blocks_per_sm = nb < 1 ? 1 : nb;
    if ((e = decoder_tc_prepare_weights(c)) != cudaSuccess) return e;
    d->version += 1;
    decoder_adopt_shared(c);
    return cudaSuccess;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

cudaError_t launch_greedy_decode(Ctx *c, const float *enc_dev, const float *enc_host, int B, int T, const int32_t *lens_dev,
                                 const int32_t *lens_host, const int32_t *slots_dev, float *s1_dev, float *s2_dev,
                                 int32_t *tokens_dev, int32_t *ntok_dev, int32_t *nsteps_dev, const int64_t *enc_off_host, int32_t *last_dev) {
    // decode_engine: 1 = fp32 CUDA-core persistent kernel (this file, the numerics anchor); 0 = auto = 4 = tcgen05 split-bf16
    // weight-stationary dataflow kernel (decoder_ws.cu, launched from decoder_tc.cu)
    // The non-reference decode rules of SURVEY 8(f4) (canonical state rule, TDT durations) run on the fp32 engine.
    if (c->cfg.decode_engine != 1 && c->cfg.decode_rule == 0)
        return launch_greedy_decode_tc(c, enc_dev, enc_host, B, T, lens_dev, lens_host, slots_dev, s1_dev, s2_dev, tokens_dev, ntok_dev,
                                       nsteps_dev, enc_off_host, last_dev);
    if (enc_off_host || last_dev) return cudaErrorNotSupported;  // packed encoder outputs: tcgen05 engines only
    DecoderPriv *d = c->dec;
    if (enc_host) {  // fp32 reference engine: plain upload
        cudaError_t eu = cudaMemcpyAsync(const_cast<float *>(enc_dev), enc_host, sizeof(float) * (size_t)B * kEnc * T, cudaMemcpyHostToDevice, c->stream);
        if (eu != cudaSuccess) return eu;
    }
    const BlobLayout L = blob_layout();
    const size_t BH = (size_t)B * kH;
    const int Tq = T > 0 ? T : 1;
    // workspace carve-up
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    const size_t oE = take(sizeof(float) * (size_t)B * Tq * kH);
    const size_t oh0 = take(sizeof(float) * 2 * BH), oh1 = take(sizeof(float) * 2 * BH);
    const size_t oc0 = take(sizeof(float) * 2 * BH), oc1 = take(sizeof(float) * 2 * BH), oz = take(sizeof(float) * BH);
    const size_t odur = take(sizeof(float) * 8 * (size_t)B);
    const size_t opv = take(sizeof(float) * (size_t)B * NT_OUT), opi = take(sizeof(int) * (size_t)B * NT_OUT);
    const size_t octl = take(sizeof(Ctl) * 2 * (size_t)B), oact = take(sizeof(int) * 4);
    cudaError_t e;
    if ((e = d->work.reserve(off)) != cudaSuccess) return e;
    char *base = d->work.as<char>();
    float *E = reinterpret_cast<float *>(base + oE);
    if ((e = cudaMemsetAsync(base + oact, 0, sizeof(int) * 4, c->stream)) != cudaSuccess) return e;

    // hoisted encoder projection: E[b][t][:] = W_enc enc[b][:, t] + b_enc + b_pred, frames < len only
    if (T > 0) {
        GemmArgs g{};
        g.A = enc_dev; g.a_batch = (size_t)kEnc * T; g.lda = T; g.a_kmajor = 1;
        g.W = c->w_blob + L.w_enc; g.ldw = kEnc; g.bias = d->bjoint; g.bias2 = nullptr;
        g.C = E; g.c_batch = (size_t)T * kH; g.ldc = kH;
        g.M = T; g.N = kH; g.K = kEnc; g.m_valid = lens_dev;
        ProfScope prof(c, PK_ENC_PROJ);
        if ((e = run_gemm(c, g, B)) != cudaSuccess) return e;
    }

    DecArgs a{};
    a.w.g0p = d->g0p; a.w.whh0p = d->whh0p; a.w.w1p = d->w1p; a.w.b1p = d->b1p;
    a.w.wpred = c->w_blob + L.w_pred; a.w.woutp = d->woutp; a.w.boutp = d->boutp;
    a.E = E; a.B = B; a.T = Tq; a.lens = lens_dev; a.slots = slots_dev;
    a.h0 = reinterpret_cast<float *>(base + oh0); a.h1 = reinterpret_cast<float *>(base + oh1);
    a.c0 = reinterpret_cast<float *>(base + oc0); a.c1 = reinterpret_cast<float *>(base + oc1);
    a.z = reinterpret_cast<float *>(base + oz);
    a.pval = reinterpret_cast<float *>(base + opv); a.pidx = reinterpret_cast<int *>(base + opi);
    a.ctl = reinterpret_cast<Ctl *>(base + octl); a.act_count = reinterpret_cast<int *>(base + oact);
    if (slots_dev) { a.s1 = c->slot_s1; a.s2 = c->slot_s2; } else { a.s1 = s1_dev; a.s2 = s2_dev; }
    a.tokens = tokens_dev; a.ntok = ntok_dev; a.nsteps = nsteps_dev;
    a.max_sym = c->cfg.max_symbols_per_step; a.max_total = c->cfg.max_total_tokens; a.blank = c->cfg.blank_id;
    a.relu = c->cfg.joint_activation;
    a.rule = c->cfg.decode_rule;
    a.dur = reinterpret_cast<float *>(base + odur);

    const int MT = (B + TM - 1) / TM;
    int grid = std::min(d->coop_blocks_per_sm * c->sm_count, MT * NT_GATES);
    if (grid < 1) grid = 1;
    d->fail_count_dev = a.act_count + 2;
    void *params[] = {&a};
    ProfScope prof(c, PK_GREEDY);
    e = cudaLaunchCooperativeKernel((const void *)greedy_persistent_kernel, dim3(grid), dim3(DEC_THREADS), params, 0, c->stream);
    c->launches++;
    return e;
}

cudaError_t launch_decoder_joint(Ctx *c, const float *enc_dev, int B, int T, const int32_t *targets_dev, int U,
                                 const int32_t *tlen_dev, const float *in_s1, const float *in_s2, float *outputs,
                                 int32_t *prednet_lengths, float *out_s1, float *out_s2, int32_t *err_flag_dev) {
    DecoderPriv *d = c->dec;
    const BlobLayout L = blob_layout();
    const size_t BH = (size_t)B * kH;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    const size_t oE = take(sizeof(float) * (size_t)B * T * kH);
    const size_t oh0 = take(sizeof(float) * 2 * BH), oh1 = take(sizeof(float) * 2 * BH);
    const size_t oc = take(sizeof(float) * 2 * BH);
    const size_t oP = take(sizeof(float) * (size_t)B * U * kH);
    const size_t oz = take(sizeof(float) * (size_t)B * U * T * kH);
    cudaError_t e;
    if ((e = d->work.reserve(off)) != cudaSuccess) return e;
    char *base = d->work.as<char>();
    float *E = reinterpret_cast<float *>(base + oE);
    float *h0 = reinterpret_cast<float *>(base + oh0), *h1 = reinterpret_cast<float *>(base + oh1);
    float *cc = reinterpret_cast<float *>(base + oc);
    float *P = reinterpret_cast<float *>(base + oP), *z = reinterpret_cast<float *>(base + oz);
    // states: [2][B][640] layer-major == {layer0 [B][640], layer1 [B][640]}
    if (in_s1) {
        if ((e = cudaMemcpyAsync(h0, in_s1, sizeof(float) * BH, cudaMemcpyDeviceToDevice, c->stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(h1, in_s1 + BH, sizeof(float) * BH, cudaMemcpyDeviceToDevice, c->stream)) != cudaSuccess) return e;
    } else {
        cudaMemsetAsync(h0, 0, sizeof(float) * BH, c->stream);
        cudaMemsetAsync(h1, 0, sizeof(float) * BH, c->stream);
    }
    if (in_s2) {
        if ((e = cudaMemcpyAsync(cc, in_s2, sizeof(float) * 2 * BH, cudaMemcpyDeviceToDevice, c->stream)) != cudaSuccess) return e;
    } else {
        cudaMemsetAsync(cc, 0, sizeof(float) * 2 * BH, c->stream);
    }
    {
        GemmArgs g{};
        g.A = enc_dev; g.a_batch = (size_t)kEnc * T; g.lda = T; g.a_kmajor = 1;
        g.W = c->w_blob + L.w_enc; g.ldw = kEnc; g.bias = d->bjoint; g.bias2 = nullptr;
        g.C = E; g.c_batch = (size_t)T * kH; g.ldc = kH;
        g.M = T; g.N = kH; g.K = kEnc; g.m_valid = nullptr;
        if ((e = run_gemm(c, g, B)) != cudaSuccess) return e;
    }
    const int MT = (B + TM - 1) / TM;
    int par = 0;
    for (int u = 0; u < U; ++u) {
        StepArgs q{};
        q.B = B; q.U = U; q.u = u; q.targets = targets_dev; q.tlen = tlen_dev; q.err_flag = err_flag_dev;
        // layer 0
        q.W = d->whh0p; q.ldw = kH; q.nseg = 1; q.a0 = h0 + (size_t)par * BH; q.a1 = nullptr; q.g0p = d->g0p; q.bias = nullptr;
        q.c = cc; q.h_old = h0 + (size_t)par * BH; q.h_new = h0 + (size_t)(par ^ 1) * BH;
        lstm_step_kernel<<<dim3(NT_GATES, MT), DEC_THREADS, 0, c->stream>>>(q);
        // layer 1
        q.W = d->w1p; q.ldw = 2 * kH; q.nseg = 2; q.a0 = h0 + (size_t)(par ^ 1) * BH; q.a1 = h1 + (size_t)par * BH; q.g0p = nullptr;
        q.bias = d->b1p; q.c = cc + BH; q.h_old = h1 + (size_t)par * BH; q.h_new = h1 + (size_t)(par ^ 1) * BH;
        lstm_step_kernel<<<dim3(NT_GATES, MT), DEC_THREADS, 0, c->stream>>>(q);
        c->launches += 2;
        par ^= 1;
        // prediction projection for position u: P[b][u][:] = W_pred h1'
        GemmArgs g{};
        g.A = h1 + (size_t)par * BH; g.a_batch = 0; g.lda = kH; g.a_kmajor = 0;
        g.W = c->w_blob + L.w_pred; g.ldw = kH; g.bias = nullptr; g.bias2 = nullptr;
        g.C = P + (size_t)u * kH; g.c_batch = 0; g.ldc = U * kH;
        g.M = B; g.N = kH; g.K = kH; g.m_valid = nullptr;
        if ((e = run_gemm(c, g, 1)) != cudaSuccess) return e;
    }
    {
        const size_t n = (size_t)B * U * T * kH;
        const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 16);
        joint_hidden_kernel<<<blocks, 256, 0, c->stream>>>(E, P, z, B, U, T, c->cfg.joint_activation);
        c->launches++;
        GemmArgs g{};
        g.A = z; g.a_batch = 0; g.lda = kH; g.a_kmajor = 0;
        g.W = c->w_blob + L.w_out; g.ldw = kH; g.bias = c->w_blob + L.b_out; g.bias2 = nullptr;
        g.C = outputs; g.c_batch = 0; g.ldc = kV;
        g.M = B * U * T; g.N = kV; g.K = kH; g.m_valid = nullptr;
        if ((e = run_gemm(c, g, 1)) != cudaSuccess) return e;
    }
    if (out_s1) {
        cudaMemcpyAsync(out_s1, h0 + (size_t)par * BH, sizeof(float) * BH, cudaMemcpyDeviceToDevice, c->stream);
        cudaMemcpyAsync(out_s1 + BH, h1 + (size_t)par * BH, sizeof(float) * BH, cudaMemcpyDeviceToDevice, c->stream);
    }
    if (out_s2) cudaMemcpyAsync(out_s2, cc, sizeof(float) * 2 * BH, cudaMemcpyDeviceToDevice, c->stream);
    if (prednet_lengths) {
        fill_int_kernel<<<(B + 255) / 256, 256, 0, c->stream>>>(prednet_lengths, tlen_dev, U, B);
        c->launches++;
    }
    return cudaGetLastError();
}

}  // namespace amira
