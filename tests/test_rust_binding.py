"""The Rust side of the boundary ships as files (no rustc in this image, so they cannot be compiled here): these CPU tests keep
them mechanically consistent with the C ABI they bind.

  rust/amira-b200-sys/src/lib.rs   GENERATED from include/amira_b200.h (scripts/gen_rust_ffi.py): every exported symbol, argument
                                   count and pointer constness; replaces src/cuda/mod.rs:371-412 of the reference
  rust/amira-b200-sys/build.rs     same translation units and nvcc flags as amira-rust-asr-server_b200/build.py; replaces
                                   build.rs:11-114 of the reference
  rust/server-patch/src/b200/      impl AsrPipeline (src/asr/pipeline.rs:20-67): all four methods, limits from src/config.rs:341-346
"""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_ffi_declarations_are_generated_from_the_header():
    gen = _load(os.path.join(ROOT, "scripts", "gen_rust_ffi.py"), "gen_rust_ffi")
    header = open(gen.HEADER).read()
    committed = open(gen.OUT).read()
    assert gen.generate(header) == committed, "rust/amira-b200-sys/src/lib.rs is stale: run scripts/gen_rust_ffi.py"


def test_every_exported_symbol_is_bound_with_the_same_arity(amira):
    gen = _load(os.path.join(ROOT, "scripts", "gen_rust_ffi.py"), "gen_rust_ffi")
    protos = {name: params for name, _, params in gen.parse_prototypes(open(gen.HEADER).read())}
    assert set(protos) == set(amira.EXPORTS)  # the header, the ctypes face and the Rust crate name the same symbols
    rust = open(gen.OUT).read()
    for name, params in protos.items():
        m = re.search(r"pub fn " + name + r"\((.*?)\) -> (i32|\*const c_char);", rust)
        assert m, name
        n_rust = 0 if not m.group(1).strip() else len(m.group(1).split(", "))
        assert n_rust == len(params), (name, n_rust, len(params))
        for (pname, ctype), rarg in zip(params, m.group(1).split(", ")):
            if "*" in ctype and not ctype.startswith("amira_encoder_fn"):
                assert ("*const" in rarg) == ctype.strip().startswith("const"), (name, pname, ctype, rarg)


def test_struct_layouts_match():
    rust = open(os.path.join(ROOT, "rust", "amira-b200-sys", "src", "lib.rs")).read()
    cfg = re.search(r"pub struct AmiraConfig \{(.*?)\}", rust, flags=re.S).group(1)
    assert re.findall(r"pub (\w+): i32", cfg) == ["device_id", "max_symbols_per_step", "max_total_tokens", "blank_id", "joint_activation",
                                                  "decode_engine", "max_streams", "decode_rule"]
    tr = re.search(r"pub struct AmiraTranscription \{(.*?)\}", rust, flags=re.S).group(1)
    assert re.findall(r"pub (\w+): (\w+)", tr) == [("audio_length_samples", "i64"), ("features_length", "i64"), ("encoded_length", "i64"),
                                                   ("n_tokens", "i32"), ("text_len", "i32")]


def test_build_rs_compiles_the_same_sources_with_the_same_arch():
    b = _load(os.path.join(ROOT, "amira-rust-asr-server_b200", "build.py"), "amira_build_py")
    rs = open(os.path.join(ROOT, "rust", "amira-b200-sys", "build.rs")).read()
    srcs = re.findall(r'"([a-z_]+\.(?:cu|cpp))"', re.search(r"const SOURCES: &\[&str\] = &\[(.*?)\];", rs, flags=re.S).group(1))
    assert srcs == b.SOURCES
    for s in srcs:
        assert os.path.exists(os.path.join(ROOT, "amira-rust-asr-server_b200", "csrc", s)), s
    assert "arch=compute_100a,code=sm_100a" in rs and "rustc-link-lib=dylib=amira_b200" in rs
    assert "tritonserver" not in rs  # the Triton link lines of the reference's build.rs:54-60 are gone on this path


def test_pipeline_implements_the_four_trait_methods_and_wires_the_config_limits():
    p = open(os.path.join(ROOT, "rust", "server-patch", "src", "b200", "pipeline.rs")).read()
    assert "impl AsrPipeline for B200AsrPipeline" in p
    for sig in ("async fn process_stream_chunk(&self, audio_bytes: &[u8], state: &mut DecoderState) -> Result<Transcription>",
                "async fn process_batch(&self, audio_bytes: &[u8]) -> Result<Transcription>",
                "async fn process_stream_samples(&self, audio_samples: &[f32], state: &mut DecoderState) -> Result<Transcription>",
                "async fn process_batch_samples(&self, audio_samples: &[f32]) -> Result<Transcription>"):
        assert sig in p, sig
    assert "config.max_symbols_per_step" in p and "config.max_total_tokens" in p and "config.cuda_device_id" in p
    safe = open(os.path.join(ROOT, "rust", "amira-b200-sys", "src", "safe.rs")).read()
    rust = open(os.path.join(ROOT, "rust", "amira-b200-sys", "src", "lib.rs")).read()
    for sym in set(re.findall(r"\b(amira_[a-z0-9_]+)\(", safe)):
        assert f"pub fn {sym}(" in rust, sym  # the safe layer only calls functions the header declares
    main_patch = open(os.path.join(ROOT, "rust", "server-patch", "main.rs.patch")).read()
    assert "is_b200_backend" in main_patch and "B200AsrPipeline::new" in main_patch
