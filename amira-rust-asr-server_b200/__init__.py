"""amira_b200 — host-side bindings over libamira_b200.so (the B200-native `preprocessor` front end and
`decoder_joint` greedy decode of amira-rust-asr-server).

The compute path is the C-ABI shared library declared in include/amira_b200.h; this module only marshals numpy
arrays / raw device pointers into it.  There is NO CPU fallback: loading fails loudly when the library is missing
and every compute entry raises AmiraError(AMIRA_ERR_NO_DEVICE) without an sm_100 GPU.

Names mirror the reference's interface (citations relative to the reference root):
  Context.preprocessor(...)   <-> PreprocessorModel (src/triton/model.rs:67-259; model-repo/preprocessor/config.pbtxt)
  Context.decoder_joint(...)  <-> DecoderJointModel (src/triton/model.rs:421-723; model-repo/decoder_joint/config.pbtxt)
  Context.greedy_decode(...)  <-> greedy_decode (src/asr/decoder_optimized.rs:24-200)
  bytes_to_f32                <-> performance_opts::audio::bytes_to_f32_optimized (src/performance_opts.rs:14-31)
  DecoderState                <-> src/asr/types.rs:159-183
"""
from __future__ import annotations

from ._lib import (AMIRA_N_PARAMS, BLANK_ID, ENC_DIM, MAX_SYMBOLS_PER_STEP, MAX_TOTAL_TOKENS, N_MELS, STATE_SIZE,
                   VOCAB_SIZE, AmiraError, Context, DecoderState, EXPORTS, device_count, device_reset, features_len, lib_path,
                   load_library, random_weights, synthetic_weights, blob_views, SYNTH_SCALE, SYNTH_BLANK_BIAS)
from .pipeline import B200AsrPipeline, Batcher, NativeEncoder, Transcription, Vocabulary, shard_utterances
from . import streaming
from .streaming import IncrementalAsr, StreamGroup

__all__ = ["NativeEncoder", "AMIRA_N_PARAMS", "BLANK_ID", "ENC_DIM", "MAX_SYMBOLS_PER_STEP", "MAX_TOTAL_TOKENS", "N_MELS", "STATE_SIZE",
           "VOCAB_SIZE", "AmiraError", "Context", "DecoderState", "EXPORTS", "device_count", "device_reset", "features_len", "lib_path",
           "load_library", "random_weights", "synthetic_weights", "blob_views", "SYNTH_SCALE", "SYNTH_BLANK_BIAS", "B200AsrPipeline", "Batcher", "Transcription", "Vocabulary", "shard_utterances", "streaming", "IncrementalAsr", "StreamGroup"]
