//! `impl AsrPipeline for B200AsrPipeline` — the drop-in for `TritonAsrPipeline` (src/asr/pipeline.rs:85-443) with two of its three
//! Triton round trips replaced by libamira_b200.so:
//!
//!   step 0/1  convert_audio + PreprocessorModel::infer_zero_copy (src/asr/pipeline.rs:127-139, 283-291)  -> Ctx::preprocess_*
//!   step 2    EncoderModel::infer (src/asr/pipeline.rs:298-312)                                          -> unchanged (Triton)
//!   step 3    greedy_decode + one DecoderJointModel::infer per step (src/asr/pipeline.rs:313-356,
//!             src/asr/decoder_optimized.rs:24-200)                                                       -> Ctx::greedy_decode
//!   step 4    Vocabulary::decode_tokens (src/asr/pipeline.rs:361-363)                                    -> unchanged
//!
//! Same trait, same `Transcription` fields, same `DecoderState` semantics (LSTM state carried across calls, token history per call).
//! Not compiled in the image this was written in (no Rust toolchain there); the C ABI it binds is exercised by the Python and
//! C++ hosts of the library repository.
use std::sync::Arc;

use async_trait::async_trait;
use tracing::{debug, info};

use amira_b200_sys::safe::{B200Error, Ctx, DecodeLimits};

use crate::asr::pipeline::AsrPipeline;
use crate::asr::types::{DecoderState, Transcription, Vocabulary};
use crate::config::Config;
use crate::error::{AppError, AsrError, Result};
use crate::triton::{ConnectionPool, EncoderInput, EncoderModel, TritonModel};

/// Lanes per GPU: independent submission queues that share one copy of the weights (amira_ctx_fork), so the upload of one
/// request overlaps the decode kernel of another.  Requests pick a lane round-robin.
const LANES: usize = 4;

pub struct B200AsrPipeline {
    lanes: Vec<Arc<Ctx>>,
    next_lane: std::sync::atomic::AtomicUsize,
    connection_pool: Arc<ConnectionPool>,
    vocabulary: Arc<Vocabulary>,
    encoder: EncoderModel,
}

fn map_err(e: B200Error) -> AppError {
    // the reference maps its CUDA FFI failures to AppError::Cuda(CudaError::Device(..)) (src/asr/cuda_pipeline.rs:55-61); that
    // variant is gated on the `cuda` feature, so the B200 backend reports through the ASR error family instead
    AppError::Asr(AsrError::Pipeline(e.to_string()))
}

impl B200AsrPipeline {
    /// `config.cuda_device_id` selects the GPU (src/config.rs:284-290); the decode limits come from the configuration
    /// (src/config.rs:341-346) instead of the compile-time constants of src/constants.rs:135-136.
    pub fn new(config: &Config, weights_path: &str, connection_pool: Arc<ConnectionPool>, vocabulary: Arc<Vocabulary>) -> Result<Self> {
        let limits = DecodeLimits { max_symbols_per_step: config.max_symbols_per_step, max_total_tokens: config.max_total_tokens };
        let first = Ctx::new(config.cuda_device_id, limits, config.max_concurrent_streams).map_err(map_err)?;
        first.load_weights_file(weights_path).map_err(map_err)?;
        let mut lanes = vec![Arc::new(first)];
        for _ in 1..LANES {
            let lane = lanes[0].fork().map_err(map_err)?;
            lanes.push(Arc::new(lane));
        }
        info!("B200 backend ready: device {}, {} lanes, limits {}/{}", config.cuda_device_id, LANES, limits.max_symbols_per_step, limits.max_total_tokens);
        Ok(Self { lanes, next_lane: Default::default(), connection_pool, vocabulary, encoder: EncoderModel })
    }

    fn lane(&self) -> Arc<Ctx> {
        let i = self.next_lane.fetch_add(1, std::sync::atomic::Ordering::Relaxed);
        self.lanes[i % self.lanes.len()].clone()
    }

    /// process_audio_zero_copy (src/asr/pipeline.rs:269-380) with the front end and the decode loop on the GPU.
    async fn run(&self, front: Front<'_>, state: &mut DecoderState) -> Result<Transcription> {
        let ctx = self.lane();
        let n_samples = match &front {
            Front::Bytes(b) => b.len() / 2,
            Front::Samples(s) => s.len(),
        };
        // step 0/1: blocking FFI off the async executor (the reference polls cudaStreamQuery instead, src/cuda/async_stream.rs:344-384)
        let (features, features_len) = tokio::task::block_in_place(|| match &front {
            Front::Bytes(b) => ctx.preprocess_pcm16(b),
            Front::Samples(s) => ctx.preprocess_f32(s),
        })
        .map_err(map_err)?;
        debug!("front end complete: features_len={}", features_len);

        // step 2: encoder, unchanged
        let encoder_output = {
            let mut pooled = self.connection_pool.get().await?;
            let mut connection = pooled.client_mut().client_mut().await;
            self.encoder.infer(&mut connection, EncoderInput { features, features_len }).await?
        };

        // step 3: the whole greedy loop in one persistent kernel; state in and out as in src/asr/pipeline.rs:392-398
        let tokens = tokio::task::block_in_place(|| {
            ctx.greedy_decode(&encoder_output.outputs, encoder_output.encoded_len, &mut state.states_1, &mut state.states_2)
        })
        .map_err(map_err)?;

        // step 4
        let text = self.vocabulary.decode_tokens(&tokens);
        Ok(Transcription {
            text,
            tokens,
            audio_length_samples: n_samples,
            features_length: features_len,
            encoded_length: encoder_output.encoded_len,
        })
    }
}

enum Front<'a> {
    Bytes(&'a [u8]),
    Samples(&'a [f32]),
}

#[async_trait]
impl AsrPipeline for B200AsrPipeline {
    // src/asr/pipeline.rs:384-401 — the LSTM state persists across chunks, the token history does not
    async fn process_stream_chunk(&self, audio_bytes: &[u8], state: &mut DecoderState) -> Result<Transcription> {
        self.run(Front::Bytes(audio_bytes), state).await
    }

    // src/asr/pipeline.rs:403-414 — fresh DecoderState
    async fn process_batch(&self, audio_bytes: &[u8]) -> Result<Transcription> {
        let mut state = DecoderState::new();
        self.run(Front::Bytes(audio_bytes), &mut state).await
    }

    // src/asr/pipeline.rs:416-431
    async fn process_stream_samples(&self, audio_samples: &[f32], state: &mut DecoderState) -> Result<Transcription> {
        self.run(Front::Samples(audio_samples), state).await
    }

    // src/asr/pipeline.rs:433-443
    async fn process_batch_samples(&self, audio_samples: &[f32]) -> Result<Transcription> {
        let mut state = DecoderState::new();
        self.run(Front::Samples(audio_samples), &mut state).await
    }
}
