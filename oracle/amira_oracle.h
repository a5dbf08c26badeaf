/*
 * amira_oracle.h — CPU ORACLE for the amira-rust-asr-server hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libamira_b200.so) never links, loads or falls back to anything in oracle/.
 *
 * It restates, in plain C, the algorithm of the two stages the reference owns around
 * its encoder.  Every function cites the reference file:line it follows (paths are
 * relative to the reference repository root).
 *
 * PARITY STATUS
 *   pinned   : i16->f32 conversion, frame gather layout, argmax tie rule, greedy loop
 *              control flow and limits — checked against every KAT the reference's own
 *              unit tests hold for this path (tests/test_oracle_kats.py lists them).
 *   UNPINNED : mel values and LSTM/joint numerics.  The arithmetic lives in two ONNX
 *              files that are Git-LFS pointers in the reference (absent):
 *                model-repo/preprocessor/1/model.onnx   sha256 c11f9e12... 140781 B
 *                model-repo/decoder_joint/1/model.onnx  sha256 cbb52a07... 35792059 B
 *              executed by Triton's onnxruntime backend (no version pinned; no Cargo.lock).
 *              The oracle follows the published NeMo AudioToMelSpectrogramPreprocessor /
 *              RNNTDecoderJoint export semantics and is cross-checked by
 *              independent CPU implementations (float64 numpy, torch.stft + torchaudio
 *              fbanks, torch.nn.LSTM) — "parity unpinned" for those values.
 */
#ifndef AMIRA_ORACLE_H
#define AMIRA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- decode constants: src/constants.rs:133-137 ---- */
#define ORC_VOCAB 1030
#define ORC_BLANK 1024
#define ORC_MAX_SYMBOLS_PER_STEP 30
#define ORC_MAX_TOTAL_TOKENS 200
#define ORC_H 640
#define ORC_ENC 1024
#define ORC_EMB_ROWS 1025
#define ORC_NMEL 128
#define ORC_NFFT 512
#define ORC_NBIN 257
#define ORC_WIN 400
#define ORC_HOP 160

/* ---- a1/a2/a3: PCM conversion ---- */
size_t orc_bytes_to_f32_optimized(const uint8_t *in, size_t n, float *out); /* performance_opts.rs:14-31 */
size_t orc_bytes_to_f32_samples(const uint8_t *in, size_t n, float *out);   /* asr/audio.rs:18-26 */
size_t orc_bytes_to_f32_simd(const uint8_t *in, size_t n, float *out);      /* asr/simd.rs:86-114,168-173,222-248 */

/* ---- a7/a9: frame gather + argmax ---- */
size_t orc_extract_frame_into(const float *data, size_t data_len, const size_t *shape, size_t ndim,
                              size_t time_step, float *out, size_t out_len); /* zero_copy.rs:49-69 */
void orc_argmax_zero_copy(const float *logits, size_t n, size_t *idx, float *val); /* zero_copy.rs:190-232 */

/* ---- a4: mel front end (spec: SURVEY.md 8c; model-repo/preprocessor/config.pbtxt:4-28) ---- */
int64_t orc_features_len(int64_t n_samples);
void orc_mel_filterbank(float *fb /* [128][257] */);
void orc_hann_window_padded(double *w512);
/* features out: [128][t_stride] (time contiguous), frames >= features_len are zero.
 * returns features_len. precision: 0 = float32 arithmetic, 1 = float64 arithmetic. */
int64_t orc_preprocess(const float *wave, int64_t n, float *features, int64_t t_stride, int precision);
/* batch helper for the CPU baseline: B utterances, OpenMP over utterances; pcm is i16 LE. */
void orc_preprocess_pcm16_batch(const int16_t *pcm, const int64_t *offsets, int B, float *features,
                                int64_t t_stride, int64_t *features_lens, int threads);

/* ---- a8: prediction net + joint (architecture pinned by the ONNX byte size, SURVEY.md finding 2) ---- */
typedef struct {
    const float *emb;                  /* [1025][640], row 1024 (blank) zero */
    const float *w_ih[2], *w_hh[2];    /* [2560][640] each, gate order i,f,g,o */
    const float *b_ih[2], *b_hh[2];    /* [2560] */
    const float *w_enc, *b_enc;        /* [640][1024], [640] */
    const float *w_pred, *b_pred;      /* [640][640], [640] */
    const float *w_out, *b_out;        /* [1030][640], [1030] */
    int act_relu;                      /* 0 = tanh (north_star), 1 = relu */
} orc_model;

#define ORC_N_PARAMS 8946310
/* lays the pointers of m over a flat blob in the documented order (DESIGN.md "weight blob"). */
void orc_model_bind(orc_model *m, const float *blob);
/* deterministic random-init blob (SplitMix64; U(-1/sqrt(H),1/sqrt(H)) for LSTM/Linear, N(0,1) embedding,
 * blank row zero); blank_bias is added to b_out[1024]. */
void orc_model_random_init(float *blob, uint64_t seed, float blank_bias);

/* Triton contract op, B = 1: model-repo/decoder_joint/config.pbtxt:4-52, src/triton/model.rs:581-722.
 * enc: [1024][T] (f-major), targets[U], states [2][1][640]; outputs [U][T][1030]; states updated in place. */
void orc_decoder_joint(const orc_model *m, const float *enc, int T, const int32_t *targets, int U,
                       float *states_1, float *states_2, float *outputs);

/* ---- a6: the greedy loop, src/asr/decoder_optimized.rs:24-200 ---- */
/* step callback = the decode_step closure of src/asr/pipeline.rs:323-348.
 * returns number of logits written (U*1030 for the real model) or <0 on error. */
typedef int (*orc_step_fn)(void *user, const float *frame, int features, const int32_t *targets, int U,
                           float *states_1, float *states_2, float *logits, int logits_cap);

typedef struct {
    int max_symbols_per_step; /* 30 */
    int max_total_tokens;     /* 200 */
    int blank;                /* 1024 */
    int single_step;          /* 1: targets=[last] (U=1, north_star); 0: literal refeed [blank]++tokens */
    /* NON-REFERENCE options, both off by default; single_step only:
     * state_update_on_nonblank_only: the canonical RNN-T rule — the prediction-net state advances only when a token is emitted
     *   (the reference replaces it after every step, blank included: src/asr/decoder_optimized.rs:154);
     * tdt_durations: Token-and-Duration-Transducer reading of the 1030 outputs the model-repo config declares — outputs
     *   [0, 1024] are token logits (1024 = blank), outputs [1025, 1029] duration logits for skips 0..4; the frame index advances
     *   by the predicted duration (a blank with duration 0 advances by 1), as in NeMo's greedy TDT decoding; compare
     *   src/triton_backends/k2_decoder/k2_decoder_backend.cc:114-253 for the reference's only other decoder. */
    int state_update_on_nonblank_only;
    int tdt_durations;
    int initial_last;         /* single_step: the LSTM input of the first step (blank for a fresh call; amira_greedy_decode_resume) */
} orc_decode_cfg;

typedef struct {
    int n_tokens;
    int n_steps;        /* step calls made */
    int frames_visited; /* encoder frames gathered */
    float min_margin;   /* smallest top1-top2 logit gap seen (real-model steps only) */
} orc_decode_stats;

/* returns 0 ok, <0 on error ("Decode step failed"). tokens cap must be >= max_total_tokens.
 * margins (nullable) receives top1-top2 per step, cap margins_cap. */
int orc_greedy_decode(const float *enc, size_t enc_len, int64_t encoded_len, float *states_1, float *states_2,
                      orc_step_fn step, void *user, const orc_decode_cfg *cfg, int32_t *tokens,
                      orc_decode_stats *stats, float *margins, int margins_cap);

/* ready-made step over the real model (user = const orc_model*). */
int orc_model_step(void *user, const float *frame, int features, const int32_t *targets, int U,
                   float *states_1, float *states_2, float *logits, int logits_cap);

/* batch helper (B independent streams, OpenMP over streams) used by the CPU baseline.
 * enc [B][1024][T], enc_lens[B]; states [2][B][640] in/out (nullable -> zeros, not returned);
 * tokens [B][max_total], n_tokens[B], n_steps[B], min_margin[B] (nullable). */
int orc_greedy_decode_batch(const orc_model *m, const float *enc, int B, int T, const int64_t *enc_lens,
                            float *states_1, float *states_2, const orc_decode_cfg *cfg, int32_t *tokens,
                            int32_t *n_tokens, int32_t *n_steps, float *min_margin, int threads);

#ifdef __cplusplus
}
#endif
#endif
