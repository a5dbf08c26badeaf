//! B200 backend of amira-rust-asr-server: the `preprocessor` and `decoder_joint` stages run in-process on the GPU through
//! libamira_b200.so (crate `amira-b200-sys`); the encoder model and everything above `AsrPipeline` stay as they are.
//! Add `#[cfg(feature = "b200")] pub mod b200;` to src/lib.rs next to `pub mod cuda;`.
pub mod pipeline;

pub use pipeline::B200AsrPipeline;
