/*
 * amira_oracle.c — CPU ORACLE (test infrastructure; see amira_oracle.h for the rules).
 *
 * Plain-C restatement of the reference hot path.  Reference citations are relative to the
 * reference repository root.  Deterministic by construction: fixed accumulation order,
 * compiled with -ffp-contract=off so results do not depend on the host's FMA support.
 */
#include "amira_oracle.h"
#include "../include/amira_hann400.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * a1: performance_opts::audio::bytes_to_f32_optimized  (src/performance_opts.rs:14-31)
 *   chunks_exact(2) -> i16::from_le_bytes -> as f32 / 32768.0 ; odd trailing byte ->
 *   (byte as i16) as f32 / 128.0 (u8 zero-extended, so 0..255 -> 0..1.992).
 * ------------------------------------------------------------------------------------------ */
size_t orc_bytes_to_f32_optimized(const uint8_t *in, size_t n, float *out) {
    size_t k = 0;
    for (size_t i = 0; i + 1 < n; i += 2) {
        int16_t s = (int16_t)((uint16_t)in[i] | ((uint16_t)in[i + 1] << 8));
        out[k++] = (float)s / 32768.0f;
    }
    if (n % 2 != 0) {
        int16_t s = (int16_t)in[n - 1]; /* `input[len-1] as i16` on a u8: zero extension */
        out[k++] = (float)s / 128.0f;
    }
    return k;
}

/* a2: asr::audio::bytes_to_f32_samples (src/asr/audio.rs:18-26) — drops an odd trailing byte. */
size_t orc_bytes_to_f32_samples(const uint8_t *in, size_t n, float *out) {
    size_t k = 0;
    for (size_t i = 0; i + 1 < n; i += 2) {
        int16_t s = (int16_t)((uint16_t)in[i] | ((uint16_t)in[i + 1] << 8));
        out[k++] = (float)s / 32768.0f;
    }
    return k;
}

/* a3: asr::simd::bytes_to_f32_safe_optimized (src/asr/simd.rs:222-248) -> _safe_avx2 (:86-114):
 *   16-byte chunks: sign-extend 8 x i16 -> i32 -> f32, multiply by 1/32768 (exact power of two, hence
 *   bit-identical to the division); remainder through the scalar loop (:168-173); < 64 bytes all scalar.
 *   An odd trailing byte is dropped. */
size_t orc_bytes_to_f32_simd(const uint8_t *in, size_t n, float *out) {
    const float scale = 1.0f / 32768.0f;
    size_t k = 0, i = 0;
    if (n >= 64) {
        for (; i + 16 <= n; i += 16) {
            for (int l = 0; l < 8; ++l) {
                int16_t s = (int16_t)((uint16_t)in[i + 2 * l] | ((uint16_t)in[i + 2 * l + 1] << 8));
                int32_t w = (int32_t)s;
                out[k++] = (float)w * scale;
            }
        }
    }
    for (; i + 1 < n; i += 2) {
        int16_t s = (int16_t)((uint16_t)in[i] | ((uint16_t)in[i + 1] << 8));
        out[k++] = (float)s / 32768.0f;
    }
    return k;
}

/* ------------------------------------------------------------------------------------------
 * a7: TensorView::extract_frame_into (src/asr/zero_copy.rs:49-69)
 *   3-D only; bounds: time_step < T and out_len >= F, else 0; out[f] = data[f*T + t] when in range.
 * ------------------------------------------------------------------------------------------ */
size_t orc_extract_frame_into(const float *data, size_t data_len, const size_t *shape, size_t ndim,
                              size_t time_step, float *out, size_t out_len) {
    if (ndim != 3) return 0;
    size_t features = shape[1], time_steps = shape[2];
    if (time_step >= time_steps || out_len < features) return 0;
    for (size_t f = 0; f < features; ++f) {
        size_t idx = f * time_steps + time_step;
        if (idx < data_len) out[f] = data[idx];
    }
    return features;
}

/* a9: argmax_zero_copy (src/asr/zero_copy.rs:190-232): first maximum under strict '>', seeded with
 * element 0; empty -> (0, 0.0).  (The 4-way unrolling does not change the visiting order.) */
void orc_argmax_zero_copy(const float *logits, size_t n, size_t *idx, float *val) {
    if (n == 0) {
        *idx = 0;
        *val = 0.0f;
        return;
    }
    size_t mi = 0;
    float mv = logits[0];
    for (size_t i = 0; i < n; ++i) {
        if (logits[i] > mv) {
            mv = logits[i];
            mi = i;
        }
    }
    *idx = mi;
    *val = mv;
}

/* ------------------------------------------------------------------------------------------
 * a4: mel front end.  The reference sends the waveform to the `preprocessor` ONNX model
 * (src/triton/model.rs:71-160, model-repo/preprocessor/config.pbtxt:4-28); the model file is absent.
 * Spec restated from SURVEY.md 8(c) (NeMo AudioToMelSpectrogramPreprocessor semantics):
 *   preemph 0.97 -> reflect pad 256 -> Hann(400, symmetric) centred in 512 -> |STFT|^2, hop 160 ->
 *   128 Slaney mel filters (Slaney norm, 0..8000 Hz) -> log(x + 2^-24) -> per-feature mean /
 *   unbiased std over valid frames, /(std + 1e-5) -> frames >= len zeroed.
 * ------------------------------------------------------------------------------------------ */
int64_t orc_features_len(int64_t n) { return n <= 0 ? 0 : n / ORC_HOP + 1; }

static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

/* librosa.filters.mel(sr=16000, n_fft=512, n_mels=128, fmin=0, fmax=8000, norm="slaney"), computed in
 * double and rounded to float32 (what NeMo stores as `fb`).  fb[m][k]. */
void orc_mel_filterbank(float *fb) {
    double mel_f[ORC_NMEL + 2];
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
    for (int i = 0; i < ORC_NMEL + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (ORC_NMEL + 1));
    for (int m = 0; m < ORC_NMEL; ++m) {
        const double fd0 = mel_f[m + 1] - mel_f[m], fd1 = mel_f[m + 2] - mel_f[m + 1];
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        for (int k = 0; k < ORC_NBIN; ++k) {
            const double fr = 8000.0 * k / (ORC_NBIN - 1);
            const double lower = (fr - mel_f[m]) / fd0, upper = (mel_f[m + 2] - fr) / fd1;
            double w = lower < upper ? lower : upper;
            if (w < 0.0) w = 0.0;
            fb[m * ORC_NBIN + k] = (float)(w * enorm);
        }
    }
}

/* torch.hann_window(400, periodic=False) (float32 table, include/amira_hann400.h) placed at offset 56 of a 512
 * frame (torch.stft window padding). */
void orc_hann_window_padded(double *w512) {
    memset(w512, 0, sizeof(double) * ORC_NFFT);
    for (int n = 0; n < ORC_WIN; ++n) {
        float f;
        memcpy(&f, &AMIRA_HANN400_BITS[n], sizeof(f));
        w512[(ORC_NFFT - ORC_WIN) / 2 + n] = (double)f;
    }
}

/* reflect index (numpy/torch "reflect", repeated for very short signals) */
static int64_t reflect_idx(int64_t i, int64_t n) {
    if (n == 1) return 0;
    const int64_t p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - i;
}

#define REAL float
#define SUFFIX f32
#define REAL_LOG logf
#include "amira_oracle_frontend.inc"
#undef REAL
#undef SUFFIX
#undef REAL_LOG
#define REAL double
#define SUFFIX f64
#define REAL_LOG log
#include "amira_oracle_frontend.inc"
#undef REAL
#undef SUFFIX
#undef REAL_LOG

int64_t orc_preprocess(const float *wave, int64_t n, float *features, int64_t t_stride, int precision) {
    return precision ? orc_preprocess_f64(wave, n, features, t_stride) : orc_preprocess_f32(wave, n, features, t_stride);
}

void orc_preprocess_pcm16_batch(const int16_t *pcm, const int64_t *offsets, int B, float *features,
                                int64_t t_stride, int64_t *features_lens, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const int64_t n = offsets[b + 1] - offsets[b];
        float *wave = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        /* a1: i16 -> f32 (same arithmetic as orc_bytes_to_f32_optimized on an even-length buffer) */
        for (int64_t i = 0; i < n; ++i) wave[i] = (float)pcm[offsets[b] + i] / 32768.0f;
        features_lens[b] = orc_preprocess_f32(wave, n, features + (size_t)b * ORC_NMEL * t_stride, t_stride);
        free(wave);
    }
}

/* ------------------------------------------------------------------------------------------
 * a8: prediction net + joint.  Conventions: SURVEY.md Appendix A (PyTorch/NeMo export): gate order
 * i,f,g,o; W_ih[4H,in], W_hh[4H,H]; layer-2 input = layer-1 h; joint = W_out*act(W_enc*e + b_enc +
 * W_pred*h2 + b_pred) + b_out.
 * ------------------------------------------------------------------------------------------ */
void orc_model_bind(orc_model *m, const float *p) {
    m->emb = p;
    p += (size_t)ORC_EMB_ROWS * ORC_H;
    for (int l = 0; l < 2; ++l) {
        m->w_ih[l] = p;
        p += (size_t)4 * ORC_H * ORC_H;
        m->w_hh[l] = p;
        p += (size_t)4 * ORC_H * ORC_H;
        m->b_ih[l] = p;
        p += 4 * ORC_H;
        m->b_hh[l] = p;
        p += 4 * ORC_H;
    }
    m->w_enc = p;
    p += (size_t)ORC_H * ORC_ENC;
    m->b_enc = p;
    p += ORC_H;
    m->w_pred = p;
    p += (size_t)ORC_H * ORC_H;
    m->b_pred = p;
    p += ORC_H;
    m->w_out = p;
    p += (size_t)ORC_VOCAB * ORC_H;
    m->b_out = p;
    p += ORC_VOCAB;
    m->act_relu = 0;
}

static uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static float u01(uint64_t *s) { return (float)(splitmix64(s) >> 40) * (1.0f / 16777216.0f); }
static void fill_uniform(float *p, size_t n, uint64_t seed, float bound) {
    uint64_t s = seed;
    for (size_t i = 0; i < n; ++i) p[i] = (2.0f * u01(&s) - 1.0f) * bound;
}
static void fill_normal(float *p, size_t n, uint64_t seed) { /* Irwin-Hall(12) - 6: libm-free N(0,1) stand-in */
    uint64_t s = seed;
    for (size_t i = 0; i < n; ++i) {
        float a = 0.0f;
        for (int j = 0; j < 12; ++j) a += u01(&s);
        p[i] = a - 6.0f;
    }
}

void orc_model_random_init(float *blob, uint64_t seed, float blank_bias) {
    const float kh = 1.0f / sqrtf((float)ORC_H), ke = 1.0f / sqrtf((float)ORC_ENC);
    float *p = blob;
    uint64_t t = 0;
    fill_normal(p, (size_t)ORC_EMB_ROWS * ORC_H, seed + 1000003ULL * (++t));
    memset(p + (size_t)ORC_BLANK * ORC_H, 0, sizeof(float) * ORC_H);
    p += (size_t)ORC_EMB_ROWS * ORC_H;
    for (int l = 0; l < 2; ++l) {
        fill_uniform(p, (size_t)4 * ORC_H * ORC_H, seed + 1000003ULL * (++t), kh); p += (size_t)4 * ORC_H * ORC_H;
        fill_uniform(p, (size_t)4 * ORC_H * ORC_H, seed + 1000003ULL * (++t), kh); p += (size_t)4 * ORC_H * ORC_H;
        fill_uniform(p, 4 * ORC_H, seed + 1000003ULL * (++t), kh); p += 4 * ORC_H;
        fill_uniform(p, 4 * ORC_H, seed + 1000003ULL * (++t), kh); p += 4 * ORC_H;
    }
    fill_uniform(p, (size_t)ORC_H * ORC_ENC, seed + 1000003ULL * (++t), ke); p += (size_t)ORC_H * ORC_ENC;
    fill_uniform(p, ORC_H, seed + 1000003ULL * (++t), ke); p += ORC_H;
    fill_uniform(p, (size_t)ORC_H * ORC_H, seed + 1000003ULL * (++t), kh); p += (size_t)ORC_H * ORC_H;
    fill_uniform(p, ORC_H, seed + 1000003ULL * (++t), kh); p += ORC_H;
    fill_uniform(p, (size_t)ORC_VOCAB * ORC_H, seed + 1000003ULL * (++t), kh); p += (size_t)ORC_VOCAB * ORC_H;
    fill_uniform(p, ORC_VOCAB, seed + 1000003ULL * (++t), kh);
    p[ORC_BLANK] += blank_bias;
}

/* fixed-order fp32 dot product: 16 interleaved partial sums, then a fixed tree.  n % 16 == 0. */
static float dotf(const float *a, const float *b, int n) {
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    for (int i = 0; i < n; i += 16)
        for (int j = 0; j < 16; ++j) acc[j] += a[i + j] * b[i + j];
    for (int w = 8; w >= 1; w >>= 1)
        for (int j = 0; j < w; ++j) acc[j] += acc[j + w];
    return acc[0];
}
static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

static void lstm_cell(const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh, const float *x,
                      float *h, float *c) {
    float g[4 * ORC_H], hn[ORC_H];
    for (int j = 0; j < 4 * ORC_H; ++j)
        g[j] = (dotf(w_ih + (size_t)j * ORC_H, x, ORC_H) + b_ih[j]) + (dotf(w_hh + (size_t)j * ORC_H, h, ORC_H) + b_hh[j]);
    for (int j = 0; j < ORC_H; ++j) {
        const float ig = sigmoidf_(g[j]), fg = sigmoidf_(g[ORC_H + j]);
        const float gg = tanhf(g[2 * ORC_H + j]), og = sigmoidf_(g[3 * ORC_H + j]);
        const float cn = fg * c[j] + ig * gg;
        c[j] = cn;
        hn[j] = og * tanhf(cn);
    }
    memcpy(h, hn, sizeof(hn));
}

/* returns 0, or -1 when a target id is outside the embedding table (an ONNX Gather failure in the
 * reference => "Decode step failed", src/asr/decoder_optimized.rs:148-152). */
static int decoder_joint_impl(const orc_model *m, const float *enc, int T, int64_t enc_t_stride, const int32_t *targets,
                              int U, float *states_1, float *states_2, float *outputs) {
    float *pred = (float *)malloc(sizeof(float) * (size_t)U * ORC_H);
    float *eproj = (float *)malloc(sizeof(float) * (size_t)T * ORC_H);
    float frame[ORC_ENC], z[ORC_H];
    /* states: [2][1][640] — layer-major (src/asr/types.rs:159-183) */
    float *h0 = states_1, *h1 = states_1 + ORC_H, *c0 = states_2, *c1 = states_2 + ORC_H;
    for (int u = 0; u < U; ++u) {
        if (targets[u] < 0 || targets[u] >= ORC_EMB_ROWS) {
            free(pred);
            free(eproj);
            return -1;
        }
        const float *x = m->emb + (size_t)targets[u] * ORC_H;
        lstm_cell(m->w_ih[0], m->w_hh[0], m->b_ih[0], m->b_hh[0], x, h0, c0);
        lstm_cell(m->w_ih[1], m->w_hh[1], m->b_ih[1], m->b_hh[1], h0, h1, c1);
        for (int j = 0; j < ORC_H; ++j) pred[(size_t)u * ORC_H + j] = dotf(m->w_pred + (size_t)j * ORC_H, h1, ORC_H) + m->b_pred[j];
    }
    for (int t = 0; t < T; ++t) {
        for (int f = 0; f < ORC_ENC; ++f) frame[f] = enc[(size_t)f * enc_t_stride + t];
        for (int j = 0; j < ORC_H; ++j) eproj[(size_t)t * ORC_H + j] = dotf(m->w_enc + (size_t)j * ORC_ENC, frame, ORC_ENC) + m->b_enc[j];
    }
    for (int u = 0; u < U; ++u)
        for (int t = 0; t < T; ++t) {
            for (int j = 0; j < ORC_H; ++j) {
                const float s = eproj[(size_t)t * ORC_H + j] + pred[(size_t)u * ORC_H + j];
                z[j] = m->act_relu ? (s > 0.0f ? s : 0.0f) : tanhf(s);
            }
            float *o = outputs + ((size_t)u * T + t) * ORC_VOCAB;
            for (int v = 0; v < ORC_VOCAB; ++v) o[v] = dotf(m->w_out + (size_t)v * ORC_H, z, ORC_H) + m->b_out[v];
        }
    free(pred);
    free(eproj);
    return 0;
}

void orc_decoder_joint(const orc_model *m, const float *enc, int T, const int32_t *targets, int U,
                       float *states_1, float *states_2, float *outputs) {
    if (decoder_joint_impl(m, enc, T, T, targets, U, states_1, states_2, outputs) != 0)
        for (size_t i = 0; i < (size_t)U * T * ORC_VOCAB; ++i) outputs[i] = NAN;
}

/* the decode_step closure (src/asr/pipeline.rs:323-348): encoder_outputs [1,1024,1] = frame. */
int orc_model_step(void *user, const float *frame, int features, const int32_t *targets, int U,
                   float *states_1, float *states_2, float *logits, int logits_cap) {
    const orc_model *m = (const orc_model *)user;
    if (features != ORC_ENC || U * ORC_VOCAB > logits_cap) return -1;
    if (decoder_joint_impl(m, frame, 1, 1, targets, U, states_1, states_2, logits) != 0) return -1;
    return U * ORC_VOCAB;
}

/* ------------------------------------------------------------------------------------------
 * a6: greedy_decode (src/asr/decoder_optimized.rs:24-39) -> greedy_decode_zero_copy (:54-200).
 * Line references in the body are to decoder_optimized.rs.
 * ------------------------------------------------------------------------------------------ */
int orc_greedy_decode(const float *enc, size_t enc_len, int64_t encoded_len, float *states_1, float *states_2,
                      orc_step_fn step, void *user, const orc_decode_cfg *cfg, int32_t *tokens,
                      orc_decode_stats *stats, float *margins, int margins_cap) {
    orc_decode_stats st = {0, 0, 0, INFINITY};
    if (stats) *stats = st;
    if (encoded_len <= 0) { /* `len / encoded_len` (:35) would divide by zero; 0 frames => nothing to decode */
        return encoded_len == 0 ? 0 : -1;
    }
    const size_t features = enc_len / (size_t)encoded_len;          /* :35 */
    const size_t shape[3] = {1, features, (size_t)encoded_len};     /* :36 */
    const size_t T = shape[2];
    const int cap = (cfg->max_total_tokens + 1) * ORC_VOCAB;
    float *frame = (float *)malloc(sizeof(float) * (features ? features : 1));
    float *logits = (float *)malloc(sizeof(float) * (size_t)cap);
    int32_t *targets = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cfg->max_total_tokens + 2));
    int total = 0, rc = 0;
    if (cfg->single_step && (cfg->state_update_on_nonblank_only || cfg->tdt_durations)) {
        /* NON-REFERENCE variants (see amira_oracle.h): canonical state rule and/or TDT durations.  Same limits, same first-max
         * argmax; the token argmax of the TDT reading runs over outputs [0, blank] only. */
        float keep1[2 * ORC_H], keep2[2 * ORC_H];
        const int n_tok = cfg->tdt_durations ? cfg->blank + 1 : ORC_VOCAB;
        size_t t = 0;
        while (t < T && total < cfg->max_total_tokens) {
            if (orc_extract_frame_into(enc, enc_len, shape, 3, t, frame, features) != features) { rc = -2; break; }
            st.frames_visited++;
            int symbols = 0, advance = 1;
            for (;;) {
                symbols += 1;
                if (symbols > cfg->max_symbols_per_step) { advance = 1; break; }
                targets[0] = total > 0 ? tokens[total - 1] : cfg->initial_last;
                memcpy(keep1, states_1, sizeof(keep1));
                memcpy(keep2, states_2, sizeof(keep2));
                const int nl = step(user, frame, (int)features, targets, 1, states_1, states_2, logits, cap);
                if (nl < 0) { rc = -1; goto done; }
                st.n_steps++;
                size_t k;
                float kv;
                orc_argmax_zero_copy(logits, (size_t)(nl < n_tok ? nl : n_tok), &k, &kv);
                int skip = -1;
                float mg = INFINITY;  /* smallest top-1 / top-2 margin of the decisions taken at this step */
                {
                    const int nk = nl < n_tok ? nl : n_tok;
                    float second = -INFINITY;
                    for (int i = 0; i < nk; ++i)
                        if ((size_t)i != k && logits[i] > second) second = logits[i];
                    if (nk > 1) mg = kv - second;
                }
                if (cfg->tdt_durations && nl >= ORC_VOCAB) {
                    size_t d;
                    float dv;
                    const int nd = ORC_VOCAB - cfg->blank - 1;
                    orc_argmax_zero_copy(logits + cfg->blank + 1, (size_t)nd, &d, &dv);
                    float second = -INFINITY;
                    for (int i = 0; i < nd; ++i)
                        if ((size_t)i != d && logits[cfg->blank + 1 + i] > second) second = logits[cfg->blank + 1 + i];
                    if (nd > 1 && dv - second < mg) mg = dv - second;
                    skip = (int)d;
                    if ((int32_t)k == cfg->blank && skip == 0) skip = 1;
                }
                if (mg < st.min_margin) st.min_margin = mg;
                if (margins && st.n_steps <= margins_cap) margins[st.n_steps - 1] = mg;
                const int is_blank = (int32_t)k == cfg->blank;
                if (is_blank && cfg->state_update_on_nonblank_only) {  /* canonical: a blank leaves the prediction net where it was */
                    memcpy(states_1, keep1, sizeof(keep1));
                    memcpy(states_2, keep2, sizeof(keep2));
                }
                if (!is_blank) {
                    tokens[total++] = (int32_t)k;
                    if (total >= cfg->max_total_tokens) break;
                }
                if (skip >= 0) {  /* TDT: the duration decides; 0 = more symbols on this frame */
                    if (skip > 0) { advance = skip; break; }
                } else if (is_blank) {
                    advance = 1;
                    break;
                }
            }
            t += (size_t)advance;
        }
        goto done;
    }
    for (size_t t = 0; t < T; ++t) {                                /* :88 */
        if (total >= cfg->max_total_tokens) break;                  /* :89-95 */
        if (orc_extract_frame_into(enc, enc_len, shape, 3, t, frame, features) != features) { rc = -2; break; } /* :98-130 */
        st.frames_visited++;
        int symbols = 0;
        for (;;) {                                                  /* :132 */
            symbols += 1;
            if (symbols > cfg->max_symbols_per_step) break;         /* :133-137 */
            int U;
            if (cfg->single_step) {                                 /* north_star "one prediction-net step" */
                targets[0] = total > 0 ? tokens[total - 1] : cfg->initial_last;
                U = 1;
            } else {                                                /* :140-142 — [blank] ++ tokens of THIS call */
                targets[0] = cfg->blank;
                for (int i = 0; i < total; ++i) targets[1 + i] = tokens[i];
                U = 1 + total;
            }
            const int nl = step(user, frame, (int)features, targets, U, states_1, states_2, logits, cap); /* :145-152 */
            if (nl < 0) { rc = -1; goto done; }                     /* "Decode step failed" */
            st.n_steps++;                                           /* state already replaced in place (:154) */
            size_t k;
            float kv;
            orc_argmax_zero_copy(logits, (size_t)nl, &k, &kv);      /* :163 — over the WHOLE vector */
            if (nl > 1) {
                float second = -INFINITY;
                for (int i = 0; i < nl; ++i)
                    if ((size_t)i != k && logits[i] > second) second = logits[i];
                const float mg = kv - second;
                if (mg < st.min_margin) st.min_margin = mg;
                if (margins && st.n_steps <= margins_cap) margins[st.n_steps - 1] = mg;
            }
            if ((int32_t)k == cfg->blank) break;                    /* :171-173 */
            tokens[total++] = (int32_t)k;                           /* :176-177 */
            if (total >= cfg->max_total_tokens) break;              /* :179-182 */
        }
        if (total >= cfg->max_total_tokens) break;                  /* :186-188 */
    }
done:
    st.n_tokens = total;
    if (stats) *stats = st;
    free(frame);
    free(logits);
    free(targets);
    return rc;
}

int orc_greedy_decode_batch(const orc_model *m, const float *enc, int B, int T, const int64_t *enc_lens,
                            float *states_1, float *states_2, const orc_decode_cfg *cfg, int32_t *tokens,
                            int32_t *n_tokens, int32_t *n_steps, float *min_margin, int threads) {
    int err = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float s1[2 * ORC_H], s2[2 * ORC_H];
        for (int l = 0; l < 2; ++l)
            for (int j = 0; j < ORC_H; ++j) {
                s1[l * ORC_H + j] = states_1 ? states_1[((size_t)l * B + b) * ORC_H + j] : 0.0f;
                s2[l * ORC_H + j] = states_2 ? states_2[((size_t)l * B + b) * ORC_H + j] : 0.0f;
            }
        const int64_t L = enc_lens ? enc_lens[b] : T;
        /* the reference hands the decoder a dense [1,1024,L] tensor; a padded batch row is compacted first */
        float *dense = (float *)malloc(sizeof(float) * (size_t)ORC_ENC * (size_t)(L > 0 ? L : 1));
        for (int f = 0; f < ORC_ENC; ++f)
            for (int64_t t = 0; t < L; ++t) dense[(size_t)f * L + t] = enc[((size_t)b * ORC_ENC + f) * T + t];
        orc_decode_stats st;
        const int rc = orc_greedy_decode(dense, (size_t)ORC_ENC * (size_t)L, L, s1, s2, orc_model_step, (void *)m, cfg,
                                         tokens + (size_t)b * cfg->max_total_tokens, &st, NULL, 0);
        free(dense);
        if (rc != 0) {
#pragma omp atomic write
            err = rc;
        }
        n_tokens[b] = st.n_tokens;
        if (n_steps) n_steps[b] = st.n_steps;
        if (min_margin) min_margin[b] = st.min_margin;
        for (int l = 0; l < 2; ++l)
            for (int j = 0; j < ORC_H; ++j) {
                if (states_1) states_1[((size_t)l * B + b) * ORC_H + j] = s1[l * ORC_H + j];
                if (states_2) states_2[((size_t)l * B + b) * ORC_H + j] = s2[l * ORC_H + j];
            }
    }
    return err;
}
