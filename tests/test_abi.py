"""The drop-in boundary: libamira_b200.so loads, exports every symbol include/amira_b200.h declares (and nothing is
declared that is not exported), and refuses to compute without a GPU instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "amira_b200.h")


def _declared():
    src = open(HEADER, encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amira_[a-z0-9_]+)\s*\(", src)) - {"amira_encoder_fn"})


def test_header_and_binding_list_agree(amira):
    assert _declared() == sorted(amira.EXPORTS)


def test_every_declared_symbol_is_exported(amira):
    lib = ctypes.CDLL(amira.lib_path())
    for name in _declared():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", amira.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (amira_[a-z0-9_]+)", out))
    assert exported == set(_declared())


def test_library_does_not_link_the_oracle(amira):
    out = subprocess.run(["ldd", amira.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", amira.lib_path()], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_host_only_entries_work_without_gpu(amira):
    assert amira.features_len(160000) == 1001 and amira.features_len(0) == 0
    assert amira.device_count() >= 0
    blob = amira.random_weights(1, 0.0)
    assert blob.size == amira.AMIRA_N_PARAMS and np.isfinite(blob).all()
    v = amira.blob_views(amira.synthetic_weights(1))
    assert v["b_out"][1025:1030].max() < -50 and np.all(v["emb"][1024] == 0)


def test_no_cpu_fallback(amira):
    """Without an sm_100 device the context cannot be created: AMIRA_ERR_NO_DEVICE, never a silent CPU path."""
    if amira.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(amira.AmiraError) as e:
        amira.Context(device_id=0)
    assert e.value.code == 5 and "no CPU path" in str(e.value)


def test_status_codes_keep_reference_meaning():
    # #[repr(C)] enum CudaError {Success=0, InvalidValue=1, OutOfMemory=2, Unknown=3, NotReady=4}, src/cuda/mod.rs:54-62
    src = open(HEADER).read()
    for name, val in (("AMIRA_OK", 0), ("AMIRA_ERR_INVALID_VALUE", 1), ("AMIRA_ERR_OUT_OF_MEMORY", 2), ("AMIRA_ERR_UNKNOWN", 3),
                      ("AMIRA_ERR_NOT_READY", 4)):
        assert re.search(rf"{name}\s*=\s*{val}\b", src), name
    for name, val in (("AMIRA_VOCAB_SIZE", 1030), ("AMIRA_BLANK_ID", 1024), ("AMIRA_MAX_SYMBOLS_PER_STEP", 30),
                      ("AMIRA_MAX_TOTAL_TOKENS", 200), ("AMIRA_STATE_SIZE", 640)):  # src/constants.rs:133-137
        assert re.search(rf"#define {name} {val}\b", src), name


def test_decode_rule_is_validated_before_any_device_is_touched():
    """amira_config.decode_rule (SURVEY 8(f4) variants): unknown bits and the combination with the tcgen05 engine are rejected
    with AMIRA_ERR_INVALID_VALUE — on a box without a GPU too, because the configuration is checked first."""
    import amira_b200 as A
    for kw in (dict(decode_rule=4), dict(decode_rule=-1), dict(decode_engine=4, decode_rule=1)):
        with pytest.raises(A.AmiraError) as e:
            A.Context(device_id=0, **kw)
        assert e.value.code == 1, kw
