/*
 * amira_b200.h — C ABI of libamira_b200.so: the B200-native (sm_100a) replacement for the two stages
 * amira-rust-asr-server owns around its encoder:
 *   (1) the `preprocessor` front end  (i16 PCM -> f32, pre-emphasis, STFT, 128-mel, log, normalise)
 *   (2) the `decoder_joint` RNN-T greedy decode loop (LSTM prediction net, joint, argmax, limits)
 *
 * This header is the single source of truth for every binding (Rust FFI binding in INTEGRATION.md, ctypes in
 * amira-rust-asr-server_b200/_lib.py, the C++ host in csrc/host_pipeline.*).  It follows the FFI conventions of the reference's
 * own CUDA crate (citations relative to the reference root):
 *   - plain `extern "C"`, POD arguments, opaque handles with create/destroy pairs
 *       (src/cuda/mod.rs:371-412, src/cuda/cuda_helper.cu:63-142)
 *   - every entry returns a status code by value, 0 = success, codes 1..4 keep the reference's meaning
 *       (`#[repr(C)] enum CudaError`, src/cuda/mod.rs:54-62; C twin src/cuda/cuda_helper.cu:14-20)
 *   - never throws / aborts across the boundary (release profile is panic="abort", Cargo.toml:134)
 *   - host input slices are borrowed for the call only; results are copied into caller-owned buffers
 *       (ReadTestData, src/cuda/cuda_helper.cu:226-264)
 *
 * Pointer arguments documented "host or device" are classified at run time (cudaPointerGetAttributes): a
 * device pointer is used in place (no copy), a host pointer is staged through the context's pinned buffers.
 * There is NO CPU fallback: every compute entry fails with AMIRA_ERR_NO_DEVICE when no sm_100 GPU is present.
 */
#ifndef AMIRA_B200_H
#define AMIRA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes: 0..4 mirror CudaError (src/cuda/mod.rs:54-62), extended upward ---- */
typedef enum {
    AMIRA_OK = 0,
    AMIRA_ERR_INVALID_VALUE = 1,
    AMIRA_ERR_OUT_OF_MEMORY = 2,
    AMIRA_ERR_UNKNOWN = 3,
    AMIRA_ERR_NOT_READY = 4,
    AMIRA_ERR_NO_DEVICE = 5,      /* no CUDA device / not sm_100: the product path refuses to run */
    AMIRA_ERR_DECODE_STEP = 6,    /* "Decode step failed" (src/asr/decoder_optimized.rs:148-152) */
    AMIRA_ERR_NO_WEIGHTS = 7,     /* decode entry called before amira_ctx_load_weights* */
    AMIRA_ERR_IO = 8
} amira_status;

/* ---- model / decode constants (src/constants.rs:133-137; model-repo config.pbtxt) ---- */
#define AMIRA_VOCAB_SIZE 1030
#define AMIRA_BLANK_ID 1024
#define AMIRA_MAX_SYMBOLS_PER_STEP 30
#define AMIRA_MAX_TOTAL_TOKENS 200
#define AMIRA_STATE_SIZE 640
#define AMIRA_ENC_DIM 1024
#define AMIRA_N_MELS 128
#define AMIRA_N_PARAMS 8946310u /* Embedding(1025x640)+2xLSTM(640)+joint(1024/640->640->1030) */

typedef struct amira_ctx amira_ctx; /* opaque; thread-safe (internally serialised); one per GPU + forked lanes (amira_ctx_fork) */

typedef struct {
    int32_t device_id;            /* config `cuda_device_id` (src/config.rs:284-290) */
    int32_t max_symbols_per_step; /* 30  — the reference hard-codes it (SURVEY finding 5-v) */
    int32_t max_total_tokens;     /* 200 */
    int32_t blank_id;             /* 1024 */
    int32_t joint_activation;     /* 0 = tanh (north_star), 1 = relu */
    int32_t decode_engine;        /* 0 = auto (4); 1 = fp32 CUDA-core persistent kernel (numerics anchor); 4 = tcgen05 split-bf16
                                   * weight-stationary dataflow kernel (one 64-feature weight slice resident in the tensor
                                   * memory of each SM: 148 CTAs in pairs, cta_group::2 MMAs).  Other values are rejected. */
    int32_t max_streams;          /* resident stream-state slots for the WebSocket path (default 1024) */
    int32_t decode_rule;          /* 0 = the reference's literal loop (src/asr/decoder_optimized.rs:54-200: state carried unconditionally,
                                   * flat argmax over all 1030 outputs).  AMIRA_RULE_* bits select the NON-REFERENCE variants of
                                   * SURVEY 8(f4), what the model family was trained for: they run on the fp32 engine (decode_engine
                                   * 0 or 1; 4 is rejected) through amira_greedy_decode / amira_stream_decode. */
} amira_config;
#define AMIRA_RULE_STATE_ON_NONBLANK 1 /* canonical RNN-T: a blank step leaves the prediction-net state where it was */
#define AMIRA_RULE_TDT_DURATIONS 2     /* outputs 1025..1029 are duration logits: token = first max over [0, blank], frame advance
                                        * = first max over the durations (0 = another symbol on this frame; blank advances >= 1) */

/* replaces get_cuda_device_count_ffi (src/cuda/cuda_helper.cu:23-30, src/cuda/mod.rs:372) */
int32_t amira_device_count(int32_t *count);
int32_t amira_config_default(amira_config *cfg);
/* Recovery from a sticky device error (a kernel watchdog trap, any device-side fault): after one, every call on every context of
 * that device fails with AMIRA_ERR_UNKNOWN.  Destroy those contexts, call this (cudaDeviceReset), create new ones.  The reference
 * has no counterpart: its CUDA path never launches a kernel (src/cuda/cuda_helper.cu). */
int32_t amira_device_reset(int32_t device_id);
/* replaces CudaAsrPipeline::new + CudaSharedMemoryRegionCreate (src/asr/cuda_pipeline.rs:41-103,
 * src/cuda/cuda_helper.cu:63-110) */
int32_t amira_ctx_create(const amira_config *cfg, amira_ctx **out);
/* replaces CudaSharedMemoryRegionDestroy (src/cuda/cuda_helper.cu:113-142) */
int32_t amira_ctx_destroy(amira_ctx *ctx);
/* Another submission lane on the same GPU.  A context serialises its calls; the reference server has up to 10 + 50 requests in
 * flight (src/config.rs:104-107).  A fork has its own streams, staging and workspace and SHARES the parent's weights (one copy
 * of the model per GPU); calls on different lanes run concurrently, so the upload of one batch overlaps the kernels of another.
 * Weights loaded through any lane are seen by all of them at their next call.  Stream slots stay with the parent.  Lanes are
 * destroyed with amira_ctx_destroy in any order; the shared memory goes with the last one. */
int32_t amira_ctx_fork(amira_ctx *parent, amira_ctx **out);
/* message of the last failing call on this ctx (NULL ctx: last failing create on this thread); owned by the
 * library — replaces the Display impl of CudaSharedMemoryError (src/cuda/mod.rs:73-82) */
const char *amira_last_error(amira_ctx *ctx);
/* run subsequent work on a caller-provided cudaStream_t (NULL = the context's own stream) */
int32_t amira_ctx_set_stream(amira_ctx *ctx, void *cuda_stream);
int32_t amira_ctx_synchronize(amira_ctx *ctx);
/* number of library kernels launched so far on this ctx (bench.py's gpu_launches) */
int32_t amira_ctx_launch_count(amira_ctx *ctx, int64_t *count);

/* per-kernel device timing for bench.py's roofline: CUDA events on the context's stream around each launch.
 * kernel ids: 0 fused front end (log-mel + normalisation), 1 unused, 2 encoder projection GEMM, 3 persistent greedy-decode, 4 bytes_to_f32.
 * amira_ctx_profile(ctx, 1) resets the totals and starts recording; amira_ctx_kernel_ms reads them. */
int32_t amira_ctx_profile(amira_ctx *ctx, int32_t enable);
int32_t amira_ctx_kernel_ms(amira_ctx *ctx, int32_t kernel, double *total_ms, int64_t *launches);

/* diagnostics: C[M][N] = A[M][K] W[N][K]^T + bias (nullable) through the tcgen05 split-bf16 GEMM that serves the
 * hoisted encoder projection; host pointers; K % 8 == 0.  Unit test hook for the UMMA descriptor conventions. */
int32_t amira_debug_tc_gemm(amira_ctx *ctx, const float *A, const float *W, const float *bias, int32_t M, int32_t N,
                            int32_t K, float *C);

/* diagnostics: per-tick globaltimer stamps [n_its][8 M-tiles][32] (n_its <= 512) of the last weight-stationary greedy launch
 * made with the environment variable AMIRA_WS_TRACE set (slot = role * 6 + event; slot 30 = control update done). */
int32_t amira_debug_ws_trace(amira_ctx *ctx, int64_t *out, int32_t n_its);

/* ---- weights: stand-in for model-repo/decoder_joint/1/model.onnx (absent LFS object) ---- */
/* flat fp32 blob, order: emb[1025][640]; per layer l=0,1: w_ih[2560][640], w_hh[2560][640], b_ih[2560],
 * b_hh[2560]; w_enc[640][1024], b_enc[640]; w_pred[640][640], b_pred[640]; w_out[1030][640], b_out[1030]. */
int32_t amira_weights_random_init(float *blob, size_t n_params, uint64_t seed, float blank_bias); /* host only */
int32_t amira_ctx_load_weights(amira_ctx *ctx, const float *blob, size_t n_params);
int32_t amira_ctx_load_weights_file(amira_ctx *ctx, const char *path); /* raw little-endian fp32 blob */

/* ---- stage 1: front end ---- */
/* features_lens rule of the preprocessor model: floor(n/160)+1 (0 for n = 0) */
int32_t amira_features_len(int64_t n_samples, int64_t *features_len);
/* replaces convert_audio + PreprocessorModel::infer_zero_copy (src/asr/pipeline.rs:127-139,283-291;
 * src/triton/model.rs:71-160).  B utterances packed back to back: utterance b = pcm[offsets[b]..offsets[b+1]).
 * pcm: host or device.  offsets, features_lens: host.  features: host or device, [B][128][t_stride] fp32,
 * frames >= features_lens[b] are zero.  t_stride >= max features_len. */
int32_t amira_preprocess_pcm16(amira_ctx *ctx, const int16_t *pcm, const int64_t *offsets, int32_t B,
                               float *features, int64_t t_stride, int64_t *features_lens);
/* Same, ragged output: utterance b's features are a dense [128][features_lens[b]] block at features + feat_offsets[b]
 * (element offsets, int64[B+1], host, non-decreasing, block >= 128 * features_len) — the shape the reference hands its
 * encoder per request (features [1][128][T_b] with features.len() == 128 * features_len, src/asr/pipeline.rs:298-301,
 * src/triton/model.rs:126-141) without padding every utterance of a batch
 * to the longest one; only valid frames cross PCIe.  Elements of a block beyond 128 * features_len are unspecified. */
int32_t amira_preprocess_pcm16_packed(amira_ctx *ctx, const int16_t *pcm, const int64_t *offsets, int32_t B,
                                      float *features, const int64_t *feat_offsets, int64_t *features_lens);
/* Triton contract form (model-repo/preprocessor/config.pbtxt:4-28): waveforms [B][n_stride] fp32 (host or
 * device), waveforms_lens [B] int64 (host) -> features [B][128][t_stride], features_lens [B] (host). */
int32_t amira_preprocess_f32(amira_ctx *ctx, const float *waveforms, int64_t n_stride, const int64_t *waveforms_lens,
                             int32_t B, float *features, int64_t t_stride, int64_t *features_lens);
/* Ragged form of the float entry: utterance b = waveforms[wave_offsets[b]..wave_offsets[b+1]) (the f32 windows of the
 * streaming path, src/asr/incremental.rs:139-160), features as in amira_preprocess_pcm16_packed. */
int32_t amira_preprocess_f32_packed(amira_ctx *ctx, const float *waveforms, const int64_t *wave_offsets, int32_t B,
                                    float *features, const int64_t *feat_offsets, int64_t *features_lens);
/* Un-normalised log-mel, ragged layout as amira_preprocess_pcm16_packed: log(mel + 2^-24) of every frame — the tensor the
 * preprocessor holds before its per-feature normalisation.  Used by the incremental streaming path, which normalises with running
 * statistics because the statistics of the whole utterance do not exist yet. */
int32_t amira_logmel_pcm16_packed(amira_ctx *ctx, const int16_t *pcm, const int64_t *offsets, int32_t B, float *features,
                                  const int64_t *feat_offsets, int64_t *features_lens);
/* replaces performance_opts::audio::bytes_to_f32_optimized (src/performance_opts.rs:14-31), including the
 * odd-trailing-byte rule; drop_odd != 0 gives bytes_to_f32_samples (src/asr/audio.rs:18-26) /
 * simd::bytes_to_f32_optimized (src/asr/simd.rs:222-248).  bytes, out: host or device. */
int32_t amira_bytes_to_f32(amira_ctx *ctx, const uint8_t *bytes, size_t n_bytes, int32_t drop_odd, float *out,
                           size_t *n_out);

/* ---- stage 2: decoder_joint ---- */
/* Triton contract op (model-repo/decoder_joint/config.pbtxt:4-52; replaces DecoderJointModel::infer,
 * src/triton/model.rs:581-722).  encoder_outputs [B][1024][T]; targets [B][U] int32; target_length [B]
 * (nullable => U); input_states_1/2 [2][B][640] (nullable => zeros); outputs [B][U][T][1030];
 * prednet_lengths [B] (nullable); output_states_1/2 [2][B][640] (nullable).  All host or device. */
int32_t amira_decoder_joint(amira_ctx *ctx, const float *encoder_outputs, int32_t B, int32_t T, const int32_t *targets,
                            int32_t U, const int32_t *target_length, const float *input_states_1,
                            const float *input_states_2, float *outputs, int32_t *prednet_lengths,
                            float *output_states_1, float *output_states_2);
/* The fused loop: replaces greedy_decode + the per-step RPC closure (src/asr/decoder_optimized.rs:24-200,
 * src/asr/pipeline.rs:313-356) for B independent utterances in one persistent kernel.
 * encoder_outputs [B][1024][T] fp32 (host or device), encoded_lengths [B] int64 host (nullable => T);
 * states_1/2 [2][B][640] in/out (host or device; nullable => zero initial state, final state dropped);
 * tokens [B][max_total_tokens] int32, n_tokens [B], n_steps [B] (nullable): host or device. */
int32_t amira_greedy_decode(amira_ctx *ctx, const float *encoder_outputs, int32_t B, int32_t T,
                            const int64_t *encoded_lengths, float *states_1, float *states_2, int32_t *tokens,
                            int32_t *n_tokens, int32_t *n_steps);
/* Same, ragged input: stream b's encoder output is a dense [1024][encoded_lengths[b]] block at
 * encoder_outputs + enc_offsets[b] (element offsets, int64[B+1], host) — the per-request tensor of
 * EncoderModel::infer (src/triton/model.rs:298-420, outputs [1][1024][T_b]) as it is, instead of a batch padded to
 * the longest stream.  decode_engine 0 / 4. */
int32_t amira_greedy_decode_packed(amira_ctx *ctx, const float *encoder_outputs, const int64_t *enc_offsets, int32_t B,
                                   const int64_t *encoded_lengths, float *states_1, float *states_2, int32_t *tokens,
                                   int32_t *n_tokens, int32_t *n_steps);

/* The loop resumed where an earlier call left a stream: besides the LSTM state, the token emitted last is carried (last_tokens
 * [B] in/out, host or device; AMIRA_BLANK_ID for a fresh stream), so decoding a stream chunk by chunk emits exactly the tokens of
 * one call over the concatenated frames (up to max_total_tokens per call).  The reference carries the state only — every call
 * starts from blank (src/asr/decoder_optimized.rs:78) — and re-decodes overlapping windows instead (src/asr/incremental.rs:136-171);
 * this entry is the "carry state + last token" decoder of SURVEY 8 f1.  Ragged input as amira_greedy_decode_packed. */
int32_t amira_greedy_decode_resume(amira_ctx *ctx, const float *encoder_outputs, const int64_t *enc_offsets, int32_t B,
                                   const int64_t *encoded_lengths, float *states_1, float *states_2, int32_t *last_tokens,
                                   int32_t *tokens, int32_t *n_tokens, int32_t *n_steps);

/* ---- WebSocket path: device-resident per-stream LSTM state (replaces the DecoderState carried by
 * IncrementalAsr, src/asr/incremental.rs:45-51,101-105, and process_stream_* in src/asr/pipeline.rs:384-431) */
int32_t amira_stream_open(amira_ctx *ctx, int32_t *slot);  /* state zeroed (DecoderState::new, types.rs:169-174) */
int32_t amira_stream_close(amira_ctx *ctx, int32_t slot);
int32_t amira_stream_get_state(amira_ctx *ctx, int32_t slot, float *states_1, float *states_2); /* [2][1][640] */
int32_t amira_stream_set_state(amira_ctx *ctx, int32_t slot, const float *states_1, const float *states_2);
/* one tick: n streams, each with an encoder chunk [n][1024][T]; state read from / written back to the slots. */
int32_t amira_stream_decode(amira_ctx *ctx, const int32_t *slots, int32_t n, const float *encoder_outputs, int32_t T,
                            const int64_t *encoded_lengths, int32_t *tokens, int32_t *n_tokens, int32_t *n_steps);

/* ---- host pipeline: C++ mirror of `trait AsrPipeline` (src/asr/pipeline.rs:20-67) and of TritonAsrPipeline's
 * process_audio_zero_copy (src/asr/pipeline.rs:269-380) with the preprocessor and decoder_joint stages on the GPU.
 * The encoder model is out of scope and stays an injected dependency: the callback receives features
 * [1][128][features_len] (host) and must return encoder outputs [1][1024][encoded_len] (host, layout f*T+t,
 * src/asr/zero_copy.rs:61-62) in a buffer it owns until the next call on the same pipeline. ---- */
typedef struct amira_pipeline amira_pipeline;
typedef int32_t (*amira_encoder_fn)(void *user, const float *features, int64_t features_len, const float **encoder_outputs,
                                    int64_t *encoded_len); /* 0 = ok; anything else fails the request */
/* mirror of `struct Transcription` (src/asr/types.rs:217-232); tokens/text go to caller-owned buffers */
typedef struct {
    int64_t audio_length_samples;
    int64_t features_length;
    int64_t encoded_length;
    int32_t n_tokens;
    int32_t text_len; /* bytes of UTF-8 text (without NUL); truncated to text_cap - 1 */
} amira_transcription;
/* vocab_path: `<token> <id>` per line (Vocabulary::load_from_file, src/asr/types.rs:87-108) */
int32_t amira_pipeline_create(amira_ctx *ctx, const char *vocab_path, amira_encoder_fn encoder, void *encoder_user,
                              amira_pipeline **out);
int32_t amira_pipeline_destroy(amira_pipeline *p);
const char *amira_pipeline_last_error(amira_pipeline *p);
/* AsrPipeline::process_batch (src/asr/pipeline.rs:403-414): fresh DecoderState */
int32_t amira_pipeline_process_batch(amira_pipeline *p, const uint8_t *audio_bytes, size_t n_bytes,
                                     amira_transcription *out, int32_t *tokens, int32_t tokens_cap, char *text,
                                     size_t text_cap);
/* AsrPipeline::process_stream_chunk (src/asr/pipeline.rs:384-401): states_1/2 [2][1][640] in/out */
int32_t amira_pipeline_process_stream_chunk(amira_pipeline *p, const uint8_t *audio_bytes, size_t n_bytes,
                                            float *states_1, float *states_2, amira_transcription *out,
                                            int32_t *tokens, int32_t tokens_cap, char *text, size_t text_cap);
/* AsrPipeline::process_batch_samples / process_stream_samples (src/asr/pipeline.rs:416-443) */
int32_t amira_pipeline_process_batch_samples(amira_pipeline *p, const float *samples, size_t n_samples,
                                             amira_transcription *out, int32_t *tokens, int32_t tokens_cap,
                                             char *text, size_t text_cap);
int32_t amira_pipeline_process_stream_samples(amira_pipeline *p, const float *samples, size_t n_samples,
                                              float *states_1, float *states_2, amira_transcription *out,
                                              int32_t *tokens, int32_t tokens_cap, char *text, size_t text_cap);
/* Vocabulary::decode_tokens (src/asr/types.rs:111-135): unknown ids skipped, U+2581 prefix -> space, trimmed.
 * returns the full length in *text_len (may exceed text_cap - 1: output truncated). */
int32_t amira_vocab_decode(amira_pipeline *p, const int32_t *tokens, int32_t n_tokens, char *text, size_t text_cap,
                           int32_t *text_len);
/* Request micro-batcher: the reference is B = 1 per request behind semaphores (src/server/state.rs:47-61);
 * this coalesces concurrent AsrPipeline::process_batch calls into one front-end launch + one persistent decode launch.
 * amira_batcher_process_batch is blocking and thread-safe; results equal the one-by-one amira_pipeline_process_batch calls.
 * max_wait_us: coalescing window after the first queued request; max_batch: requests per launch. */
typedef struct amira_batcher amira_batcher;
int32_t amira_batcher_create(amira_pipeline *p, int32_t max_batch, int32_t max_wait_us, amira_batcher **out);
int32_t amira_batcher_destroy(amira_batcher *b);
int32_t amira_batcher_process_batch(amira_batcher *b, const uint8_t *audio_bytes, size_t n_bytes, amira_transcription *out,
                                    int32_t *tokens, int32_t tokens_cap, char *text, size_t text_cap);
int32_t amira_batcher_stats(amira_batcher *b, int64_t *n_requests, int64_t *n_batches);
/* ---- streaming orchestrator: the caller of the WebSocket path ----
 * Pure host functions, literal restatements (Rust byte/char and f32 semantics kept):
 * weave_transcript_segs / best_alignment (src/asr/weaving.rs:180-280), is_overlap_silence (src/asr/weaving.rs:285-313),
 * mean_amplitude_optimized (src/performance_opts.rs:35-60), window_sequence (src/asr/audio.rs:72-132; slices[i] =
 * {source start, source end, target start, target end}; window_size must exceed leading + trailing). */
int32_t amira_weave_transcript_segs(const char *first_seg, const char *second_seg, float percent_time_overlap,
                                    float min_alignment_score, char *out, size_t out_cap, int32_t *out_len);
int32_t amira_best_alignment(const char *first, const char *second, float percent_time_overlap, int32_t *overlap, float *score);
int32_t amira_is_overlap_silence(const float *overlap_audio, size_t n, float mean_amplitude, int32_t *silent);
int32_t amira_mean_amplitude(const float *samples, size_t n, float *mean);
int32_t amira_window_sequence(int64_t total_len, int64_t window_size, int64_t leading_context, int64_t trailing_context,
                              int64_t *slices, float *overlap_ratio, int32_t cap, int32_t *n_windows);
/* A stream group = n_streams IncrementalAsr objects (src/asr/incremental.rs:35-298; OverlappingAudioBuffer
 * src/asr/audio.rs:134-293) over one pipeline.  The reference creates one per WebSocket with chunk 2.0 s, leading 1.0 s,
 * trailing 0.5 s, capacity 10 s (src/server/stream.rs:106-119) and runs every window through its own Triton round
 * trips; amira_stream_group_process_chunks is IncrementalAsr::process_chunk for n distinct streams at once: window k of
 * every stream goes through one front-end launch, the injected encoder and one decode launch.  Per stream the result
 * equals the one-at-a-time reference flow.  status[i] (nullable) receives each stream's own code; the return value is
 * the first non-zero one. */
typedef struct amira_stream_group amira_stream_group;
int32_t amira_stream_group_create(amira_pipeline *p, int32_t n_streams, float chunk_size, float leading_context,
                                  float trailing_context, float buffer_capacity, amira_stream_group **out);
int32_t amira_stream_group_destroy(amira_stream_group *g);
const char *amira_stream_group_last_error(amira_stream_group *g);
int32_t amira_stream_group_clear(amira_stream_group *g, int32_t stream);                 /* IncrementalAsr::clear */
int32_t amira_stream_group_process_chunks(amira_stream_group *g, int32_t n, const int32_t *streams,
                                          const uint8_t *const *audio_bytes, const size_t *n_bytes, int32_t *status);
int32_t amira_stream_group_transcript(amira_stream_group *g, int32_t stream, char *text, size_t text_cap, int32_t *text_len);
int32_t amira_stream_group_tokens(amira_stream_group *g, int32_t stream, int32_t *tokens, int32_t tokens_cap, int32_t *n_tokens);
int32_t amira_stream_group_audio_length(amira_stream_group *g, int32_t stream, float *seconds);
/* IncrementalAsr::process_batch (src/asr/incremental.rs:267-292) on one stream of the group */
int32_t amira_stream_group_process_batch(amira_stream_group *g, int32_t stream, const uint8_t *audio_bytes, size_t n_bytes,
                                         amira_transcription *out, int32_t *tokens, int32_t tokens_cap, char *text,
                                         size_t text_cap);
int32_t amira_stream_group_stats(amira_stream_group *g, int64_t *n_pipeline_calls, int64_t *n_rounds);
/* INCREMENTAL mode of a stream group: every chunk costs its own frames instead of every window of a 10 s buffer.
 * Per stream the group carries (a) the audio tail the next STFT frames still need (at most two hops of left context + the
 * samples after the last complete frame), (b) running per-feature statistics (count, mean, M2) for the normalisation,
 * (c) the LSTM state and the last emitted token.  amira_stream_group_process_chunks then pushes the new audio, computes the
 * log-mel frames that have become complete (a frame is complete when its 400-sample window lies inside the received audio;
 * they are bit-identical to the frames of the whole recording), normalises them with the statistics of all frames so far, runs the
 * injected encoder on the new frames only and resumes the greedy loop (amira_greedy_decode_resume); tokens are appended, no
 * weaving takes place.  amira_stream_group_flush ends a stream: the remaining frames, with the reflect padding of the true end.
 * The literal mode above remains the parity anchor of the reference's IncrementalAsr; this mode is an extension with its own
 * definition of the normalisation (running instead of per-utterance statistics). */
int32_t amira_stream_group_set_incremental(amira_stream_group *g, int32_t enable); /* before the first audio of every stream */
int32_t amira_stream_group_flush(amira_stream_group *g, int32_t n, const int32_t *streams, int32_t *status);
/* samples received, log-mel frames emitted and encoder frames decoded so far (incremental mode) */
int32_t amira_stream_group_progress(amira_stream_group *g, int32_t stream, int64_t *samples, int64_t *frames, int64_t *encoder_frames);
/* configured max_total_tokens of a context (row stride of the tokens output of the decode entries) */
int32_t amira_ctx_max_total_tokens(amira_ctx *ctx, int32_t *value);

/* ---- wire formats either side of the path (pure host code) ----
 * WebSocket binary frames (StreamProcessor::handle_audio_chunk, src/server/stream.rs:215-281): one byte = control, otherwise
 * 16-bit PCM.  The reference names two different pairs of control bytes: the server code matches END 0xFF / KEEPALIVE 0x00
 * (src/constants.rs:243-246 via src/server/stream.rs:24-26), its configuration, README and example client say END 0x00 /
 * KEEPALIVE 0x01 (src/config.rs:95-98, examples/simple_client.rs:87).  The caller picks the dialect; SERVER is what a running
 * reference server accepts. */
#define AMIRA_WIRE_DIALECT_SERVER 0
#define AMIRA_WIRE_DIALECT_DOCUMENTED 1
#define AMIRA_FRAME_AUDIO 0
#define AMIRA_FRAME_END 1
#define AMIRA_FRAME_KEEPALIVE 2
#define AMIRA_FRAME_TOO_LARGE 3        /* > 1 MiB: "Audio chunk too large" */
#define AMIRA_FRAME_UNKNOWN_CONTROL 4  /* "Unknown control byte" */
#define AMIRA_FRAME_ODD_LENGTH 5       /* "Audio data length must be even for 16-bit PCM" */
#define AMIRA_FRAME_EMPTY 6            /* "Empty audio chunk received" */
int32_t amira_wire_classify_frame(const uint8_t *data, size_t n, int32_t dialect, int32_t *kind);
/* JSON body of POST /v2/decode/batch/{model}: `BatchRequest` + `validate()` (src/server/handlers.rs:44-116).  audio_buffer is a
 * JSON array of integers 0..255; `opaque` comes back as the span [*opaque_begin, +*opaque_len) of its raw value inside `json`
 * (length 0 = absent / null); unknown keys are ignored as serde does.  A refused request returns AMIRA_ERR_INVALID_VALUE with the
 * reference's validation message in err; audio_cap too small returns AMIRA_ERR_OUT_OF_MEMORY with the needed size in *n_audio. */
int32_t amira_wire_parse_batch_request(const char *json, size_t n, uint8_t *audio, size_t audio_cap, size_t *n_audio,
                                       size_t *opaque_begin, size_t *opaque_len, char *err, size_t err_cap);
/* `AsrResponse` as serde writes it (src/asr/types.rs:236-272: camelCase keys, None omitted, status ACTIVE / COMPLETE / PAUSED /
 * ERROR = 0..3) with the batch handler's metadata object (src/server/handlers.rs:192-206) when meta != NULL; opaque_json is
 * spliced in verbatim.  *out_len receives the full length; a short buffer truncates and returns AMIRA_ERR_OUT_OF_MEMORY. */
int32_t amira_wire_format_response(const char *transcription, int32_t status, const char *message, const amira_transcription *meta,
                                   const int32_t *tokens, const char *opaque_json, char *out, size_t out_cap, size_t *out_len);

/* ---- device-resident hand-off (replaces the CUDA shared-memory regions of src/cuda/cuda_helper.cu:63-183: cudaMalloc +
 * cudaIpcGetMemHandle, GetRawHandle).  The stages either side of the encoder exchange fp32 tensors with it (features out,
 * encoder outputs in: 51 200 + 4 096 x frames bytes per audio-second over PCIe when they go through host memory).  Every
 * compute entry already takes device pointers in place; these entries give the two processes a buffer to share:
 *   the library side allocates a region and EXPORTS it (the encoder process opens the handle and reads features from it / writes
 *   encoder outputs into it), or IMPORTS a region the encoder process exported.  The 64 bytes are a cudaIpcMemHandle_t;
 *   a handle cannot be opened in the process that created it (CUDA rule). */
typedef struct {
    uint8_t reserved[64];
} amira_ipc_handle;
int32_t amira_device_alloc(amira_ctx *ctx, size_t bytes, void **dev_ptr);      /* CudaSharedMemoryRegionCreate */
int32_t amira_device_free(amira_ctx *ctx, void *dev_ptr);                      /* CudaSharedMemoryRegionDestroy */
int32_t amira_ipc_export(amira_ctx *ctx, const void *dev_ptr, amira_ipc_handle *handle); /* GetRawHandle */
int32_t amira_ipc_import(amira_ctx *ctx, const amira_ipc_handle *handle, void **dev_ptr);
int32_t amira_ipc_close(amira_ctx *ctx, void *dev_ptr);

/* Multi-GPU: utterances are independent — longest-processing-time assignment of n utterances with
 * costs[i] (e.g. samples) to n_shards GPUs; shard_of[i] receives the shard index.  No collective follows. */
int32_t amira_shard_utterances(const int64_t *costs, int32_t n, int32_t n_shards, int32_t *shard_of);

#ifdef __cplusplus
}
#endif
#endif /* AMIRA_B200_H */
