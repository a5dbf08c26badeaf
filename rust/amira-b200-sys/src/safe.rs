//! Safe wrappers over the raw declarations in lib.rs: RAII handles, `Result` instead of status codes.
//! Mirrors the shape of the reference's wrappers around its CUDA FFI (src/cuda/mod.rs:420-760: handle structs with `Drop`,
//! `unsafe impl Send + Sync`, error enum mapped from the C status).
use std::ffi::{CStr, CString};
use std::os::raw::c_void;
use std::ptr;

use crate::*;

/// `CudaSharedMemoryError`'s role (src/cuda/mod.rs:64-96): the C status plus the library's message.
#[derive(Debug, Clone)]
pub struct B200Error {
    pub code: i32,
    pub message: String,
}

impl std::fmt::Display for B200Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "amira_b200 status {}: {}", self.code, self.message)
    }
}
impl std::error::Error for B200Error {}

pub type B200Result<T> = std::result::Result<T, B200Error>;

/// Decode limits; the reference keeps them in `Config` (src/config.rs:341-346) and in src/constants.rs:133-137.
#[derive(Debug, Clone, Copy)]
pub struct DecodeLimits {
    pub max_symbols_per_step: usize,
    pub max_total_tokens: usize,
}

impl Default for DecodeLimits {
    fn default() -> Self {
        Self { max_symbols_per_step: AMIRA_MAX_SYMBOLS_PER_STEP, max_total_tokens: AMIRA_MAX_TOTAL_TOKENS }
    }
}

/// One GPU context (`amira_ctx`).  Thread-safe: the library serialises calls on a lane; `fork()` gives another lane that
/// shares the weights, so several batches can be in flight on one GPU.
pub struct Ctx {
    raw: *mut AmiraCtx,
    max_total_tokens: usize,
}

// the handle is an opaque pointer to a mutex-protected object, like the handles of src/cuda/mod.rs:761-762
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

impl Ctx {
    pub fn new(device_id: i32, limits: DecodeLimits, max_streams: usize) -> B200Result<Self> {
        let mut cfg = AmiraConfig {
            device_id: 0, max_symbols_per_step: 0, max_total_tokens: 0, blank_id: 0, joint_activation: 0, decode_engine: 0,
            max_streams: 0, decode_rule: 0,
        };
        unsafe { amira_config_default(&mut cfg) };
        cfg.device_id = device_id;
        cfg.max_symbols_per_step = limits.max_symbols_per_step as i32;
        cfg.max_total_tokens = limits.max_total_tokens as i32;
        cfg.max_streams = max_streams as i32;
        let mut raw: *mut AmiraCtx = ptr::null_mut();
        let rc = unsafe { amira_ctx_create(&cfg, &mut raw) };
        if rc != AMIRA_OK {
            return Err(B200Error { code: rc, message: last_error(ptr::null_mut()) });
        }
        Ok(Self { raw, max_total_tokens: limits.max_total_tokens })
    }

    /// Another lane on the same GPU that shares this context's weights (amira_ctx_fork).
    pub fn fork(&self) -> B200Result<Self> {
        let mut raw: *mut AmiraCtx = ptr::null_mut();
        let rc = unsafe { amira_ctx_fork(self.raw, &mut raw) };
        self.check(rc)?;
        Ok(Self { raw, max_total_tokens: self.max_total_tokens })
    }

    pub fn raw(&self) -> *mut AmiraCtx {
        self.raw
    }

    pub fn max_total_tokens(&self) -> usize {
        self.max_total_tokens
    }

    pub fn check(&self, rc: i32) -> B200Result<()> {
        if rc == AMIRA_OK {
            Ok(())
        } else {
            Err(B200Error { code: rc, message: last_error(self.raw) })
        }
    }

    pub fn load_weights_file(&self, path: &str) -> B200Result<()> {
        let c = CString::new(path).map_err(|_| B200Error { code: AMIRA_ERR_INVALID_VALUE, message: "path contains NUL".into() })?;
        self.check(unsafe { amira_ctx_load_weights_file(self.raw, c.as_ptr()) })
    }

    /// `convert_audio` + `PreprocessorModel::infer_zero_copy` (src/asr/pipeline.rs:127-139, 283-291) for one utterance of
    /// little-endian 16-bit PCM.  Returns (features `[128][features_len]`, features_len).
    pub fn preprocess_pcm16(&self, audio_bytes: &[u8]) -> B200Result<(Vec<f32>, i64)> {
        let n = (audio_bytes.len() / 2) as i64;
        let mut flen = 0i64;
        unsafe { amira_features_len(n, &mut flen) };
        let mut feats = vec![0f32; AMIRA_N_MELS * (flen.max(1) as usize)];
        let offsets = [0i64, n];
        self.check(unsafe {
            amira_preprocess_pcm16(self.raw, audio_bytes.as_ptr() as *const i16, offsets.as_ptr(), 1, feats.as_mut_ptr(), flen.max(1), &mut flen)
        })?;
        feats.truncate(AMIRA_N_MELS * flen as usize);
        Ok((feats, flen))
    }

    /// The f32 entry of the same stage (`process_*_samples`, src/asr/pipeline.rs:416-443).
    pub fn preprocess_f32(&self, samples: &[f32]) -> B200Result<(Vec<f32>, i64)> {
        let n = samples.len() as i64;
        let mut flen = 0i64;
        unsafe { amira_features_len(n, &mut flen) };
        let mut feats = vec![0f32; AMIRA_N_MELS * (flen.max(1) as usize)];
        self.check(unsafe { amira_preprocess_f32(self.raw, samples.as_ptr(), n, &n, 1, feats.as_mut_ptr(), flen.max(1), &mut flen) })?;
        feats.truncate(AMIRA_N_MELS * flen as usize);
        Ok((feats, flen))
    }

    /// `greedy_decode` + the per-step RPC closure (src/asr/decoder_optimized.rs:24-200, src/asr/pipeline.rs:313-356) for one
    /// stream: encoder output `[1][1024][encoded_len]`, `DecoderState` (two `[2][1][640]` vectors) in and out.
    pub fn greedy_decode(&self, encoder_outputs: &[f32], encoded_len: i64, states_1: &mut [f32], states_2: &mut [f32]) -> B200Result<Vec<i32>> {
        assert_eq!(states_1.len(), 2 * AMIRA_STATE_SIZE);
        assert_eq!(states_2.len(), 2 * AMIRA_STATE_SIZE);
        assert!(encoder_outputs.len() as i64 >= AMIRA_ENC_DIM as i64 * encoded_len);
        let mut tokens = vec![0i32; self.max_total_tokens];
        let mut n_tokens = 0i32;
        if encoded_len > 0 {
            self.check(unsafe {
                amira_greedy_decode(self.raw, encoder_outputs.as_ptr(), 1, encoded_len as i32, &encoded_len, states_1.as_mut_ptr(),
                                    states_2.as_mut_ptr(), tokens.as_mut_ptr(), &mut n_tokens, ptr::null_mut())
            })?;
        }
        tokens.truncate(n_tokens.max(0) as usize);
        Ok(tokens)
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { amira_ctx_destroy(self.raw) };
    }
}

fn last_error(ctx: *mut AmiraCtx) -> String {
    let p = unsafe { amira_last_error(ctx) };
    if p.is_null() {
        return String::new();
    }
    unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
}

/// The request micro-batcher (`amira_batcher_*`): concurrent `process_batch` calls coalesce into one front-end launch and one
/// decode launch.  The encoder stays an injected dependency: a C callback that receives features `[1][128][features_len]`
/// and returns encoder outputs `[1][1024][encoded_len]`.
pub struct Batcher {
    pipeline: *mut AmiraPipeline,
    batcher: *mut AmiraBatcher,
    _ctx: std::sync::Arc<Ctx>,
}
unsafe impl Send for Batcher {}
unsafe impl Sync for Batcher {}

impl Batcher {
    /// # Safety
    /// `encoder_user` must stay valid for the life of the batcher and `encoder` must be callable from the batcher's worker thread.
    pub unsafe fn new(ctx: std::sync::Arc<Ctx>, vocab_path: &str, encoder: AmiraEncoderFn, encoder_user: *mut c_void, max_batch: i32,
                      max_wait_us: i32) -> B200Result<Self> {
        let c = CString::new(vocab_path).map_err(|_| B200Error { code: AMIRA_ERR_INVALID_VALUE, message: "path contains NUL".into() })?;
        let mut pipeline: *mut AmiraPipeline = ptr::null_mut();
        let rc = amira_pipeline_create(ctx.raw(), c.as_ptr(), encoder, encoder_user, &mut pipeline);
        if rc != AMIRA_OK {
            return Err(B200Error { code: rc, message: CStr::from_ptr(amira_pipeline_last_error(ptr::null_mut())).to_string_lossy().into_owned() });
        }
        let mut batcher: *mut AmiraBatcher = ptr::null_mut();
        let rc = amira_batcher_create(pipeline, max_batch, max_wait_us, &mut batcher);
        if rc != AMIRA_OK {
            amira_pipeline_destroy(pipeline);
            return Err(B200Error { code: rc, message: "amira_batcher_create failed".into() });
        }
        Ok(Self { pipeline, batcher, _ctx: ctx })
    }

    /// Blocking; call from `tokio::task::spawn_blocking`.  Returns (text, tokens, transcription record).
    pub fn process_batch(&self, audio_bytes: &[u8], max_total_tokens: usize) -> B200Result<(String, Vec<i32>, AmiraTranscription)> {
        let mut out = AmiraTranscription { audio_length_samples: 0, features_length: 0, encoded_length: 0, n_tokens: 0, text_len: 0 };
        let mut tokens = vec![0i32; max_total_tokens];
        let mut text = vec![0u8; 8 * max_total_tokens + 16];
        let rc = unsafe {
            amira_batcher_process_batch(self.batcher, audio_bytes.as_ptr(), audio_bytes.len(), &mut out, tokens.as_mut_ptr(), tokens.len() as i32,
                                        text.as_mut_ptr() as *mut std::os::raw::c_char, text.len())
        };
        if rc != AMIRA_OK {
            let msg = unsafe { CStr::from_ptr(amira_pipeline_last_error(self.pipeline)) }.to_string_lossy().into_owned();
            return Err(B200Error { code: rc, message: msg });
        }
        tokens.truncate(out.n_tokens.max(0) as usize);
        let len = (out.text_len.max(0) as usize).min(text.len() - 1);
        text.truncate(len);
        Ok((String::from_utf8_lossy(&text).into_owned(), tokens, out))
    }
}

impl Drop for Batcher {
    fn drop(&mut self) {
        unsafe {
            amira_batcher_destroy(self.batcher);
            amira_pipeline_destroy(self.pipeline);
        }
    }
}
