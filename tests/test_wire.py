"""Wire formats either side of the hot path through the C ABI (host code only, no GPU):
WebSocket frames (src/server/stream.rs:215-281), the JSON batch body (src/server/handlers.rs:44-116) and the AsrResponse JSON
(src/asr/types.rs:236-272).  tests/golden/wire_fixture.json holds response cases written by hand from those lines."""
import ctypes as C
import json
import os

import numpy as np
import pytest

SERVER, DOCUMENTED = 0, 1
AUDIO, END, KEEPALIVE, TOO_LARGE, UNKNOWN_CONTROL, ODD_LENGTH, EMPTY = range(7)


@pytest.fixture(scope="module")
def L(amira):
    return amira.load_library()


def classify(L, data: bytes, dialect: int) -> int:
    kind = C.c_int32(-1)
    buf = (C.c_ubyte * max(len(data), 1)).from_buffer_copy(data.ljust(1, b"\0"))
    assert L.amira_wire_classify_frame(C.cast(buf, C.c_void_p), len(data), dialect, C.byref(kind)) == 0
    return kind.value


def test_control_bytes_of_both_dialects(L):
    """The server code matches END 0xFF / KEEPALIVE 0x00 (src/constants.rs:243-246); config.rs:95-98, the README and the example
    client say END 0x00 / KEEPALIVE 0x01.  Pinned: what each byte means under each dialect, including the documented client
    against the real server (its END reads as KEEPALIVE, its KEEPALIVE is an unknown control byte)."""
    assert classify(L, b"\xff", SERVER) == END and classify(L, b"\x00", SERVER) == KEEPALIVE
    assert classify(L, b"\x00", DOCUMENTED) == END and classify(L, b"\x01", DOCUMENTED) == KEEPALIVE
    assert classify(L, b"\x01", SERVER) == UNKNOWN_CONTROL      # a documented keepalive sent to the real server
    assert classify(L, b"\xff", DOCUMENTED) == UNKNOWN_CONTROL
    assert classify(L, b"\x7f", SERVER) == UNKNOWN_CONTROL


def test_frame_validation_order(L):
    pcm = np.arange(2560, dtype=np.int16).tobytes()
    for d in (SERVER, DOCUMENTED):
        assert classify(L, pcm, d) == AUDIO
        assert classify(L, pcm[:-1], d) == ODD_LENGTH               # stream.rs:257-261
        assert classify(L, b"", d) == EMPTY                         # :264-268
        assert classify(L, b"\0" * (1024 * 1024 + 2), d) == TOO_LARGE  # :220-228, checked before everything else
        assert classify(L, b"\0" * (1024 * 1024 + 1), d) == TOO_LARGE
        assert classify(L, b"\0" * (1024 * 1024), d) == AUDIO
    kind = C.c_int32()
    assert L.amira_wire_classify_frame(None, 0, 7, C.byref(kind)) == 1  # unknown dialect


def parse(L, body: str, cap: int = 1 << 20):
    raw = body.encode()
    audio = (C.c_ubyte * cap)()
    n, ob, ol = C.c_size_t(), C.c_size_t(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = L.amira_wire_parse_batch_request(raw, len(raw), C.cast(audio, C.c_void_p), cap, C.byref(n), C.byref(ob), C.byref(ol), err, 256)
    return rc, bytes(audio[:min(n.value, cap)]), raw[ob.value:ob.value + ol.value].decode(), err.value.decode(), n.value


def test_batch_request_body(L):
    pcm = (np.arange(320) * 37 % 256).astype(np.uint8)
    body = json.dumps({"audio_buffer": pcm.tolist(), "_description": "x", "opaque": {"id": [1, 'a"b'], "n": None}, "_incremental": False,
                       "model": "amira", "extra": [[], {}]})
    rc, audio, opaque, err, n = parse(L, body)
    assert rc == 0, err
    assert audio == pcm.tobytes() and n == 320
    assert json.loads(opaque) == {"id": [1, 'a"b'], "n": None}
    rc, audio, opaque, err, _ = parse(L, ' {\n "opaque" : null ,\t"audio_buffer" : [ 255 , 0 ] } ')
    assert rc == 0 and audio == b"\xff\x00" and opaque == ""


@pytest.mark.parametrize("body,msg", [
    ('{"audio_buffer": []}', "Audio buffer cannot be empty"),
    ('{"audio_buffer": [1, 2, 3]}', "Audio buffer length must be even for 16-bit PCM"),
    ('{"audio_buffer": [1, 256]}', "integers 0..255"),
    ('{"audio_buffer": [1, -1]}', "integers 0..255"),
    ('{"audio_buffer": [1, 2.0]}', "integers 0..255"),
    ('{"audio_buffer": "AAAA"}', "must be an array"),
    ('{"opaque": 1}', "missing field `audio_buffer`"),
    ('{"audio_buffer": [1, 2], "audio_buffer": [3, 4]}', "duplicate field"),
    ('{"audio_buffer": [1, 2]} x', "trailing characters"),
    ('[1, 2]', "expected an object"),
    ('{"audio_buffer": [1, 2', "invalid JSON"),
])
def test_batch_request_refusals(L, body, msg):
    rc, _, _, err, _ = parse(L, body)
    assert rc == 1 and msg in err, err


def test_batch_request_limits(L):
    too_long = '{"audio_buffer": [' + ",".join(["0"] * (960000 + 2)) + "]}"  # 30 s of 16-bit PCM + one sample
    rc, _, _, err, _ = parse(L, too_long, cap=1 << 21)
    assert rc == 1 and err.startswith("Audio too long: 30.0s (max: 30s)")
    ok = '{"audio_buffer": [' + ",".join(["0"] * 960000) + "]}"
    rc, audio, _, err, n = parse(L, ok, cap=1 << 21)
    assert rc == 0 and n == 960000
    rc, _, _, err, n = parse(L, ok, cap=16)  # caller's buffer too small: size reported
    assert rc == 2 and n == 960000
    big_opaque = '{"audio_buffer": [0, 0], "opaque": "' + "x" * 10001 + '"}'
    rc, _, _, err, _ = parse(L, big_opaque)
    assert rc == 1 and err == "Opaque data too large (max: 10KB)"


class _Tr(C.Structure):
    _fields_ = [("audio_length_samples", C.c_int64), ("features_length", C.c_int64), ("encoded_length", C.c_int64),
                ("n_tokens", C.c_int32), ("text_len", C.c_int32)]


def fmt(L, text, status, message=None, meta=None, tokens=None, opaque=None, cap=4096):
    out = C.create_string_buffer(cap)
    n = C.c_size_t()
    tk = None if tokens is None else (C.c_int32 * max(len(tokens), 1))(*tokens)
    rc = L.amira_wire_format_response(text.encode(), status, None if message is None else message.encode(),
                                      None if meta is None else C.cast(C.pointer(meta), C.c_void_p),
                                      None if tk is None else C.cast(tk, C.c_void_p), None if opaque is None else opaque.encode(),
                                      C.cast(out, C.c_void_p), cap, C.byref(n))
    return rc, out.value.decode(), n.value


def test_response_json_matches_the_golden_fixture(L):
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wire_fixture.json")))
    for case in fx["responses"]:
        meta = None
        if case.get("metadata"):
            md = case["metadata"]
            meta = _Tr(md["audio_length_samples"], md["features_length"], md["encoded_length"], len(md["tokens"]), 0)
        rc, got, n = fmt(L, case["transcription"], ["ACTIVE", "COMPLETE", "PAUSED", "ERROR"].index(case["status"]), case.get("message"),
                         meta, case["metadata"]["tokens"] if meta else None, json.dumps(case["opaque"]) if "opaque" in case else None)
        assert rc == 0 and n == len(got.encode())
        assert got == case["wire"], case["name"]          # byte for byte what serde_json writes for AsrResponse
        back = json.loads(got)
        assert back["transcription"] == case["transcription"] and back["status"] == case["status"]
        assert ("message" in back) == ("message" in case) and ("metadata" in back) == bool(case.get("metadata"))


def test_response_truncation_and_arguments(L):
    rc, got, n = fmt(L, "hello world", 1, cap=16)
    assert rc == 2 and n > 16 and len(got) == 15
    assert L.amira_wire_format_response(b"x", 4, None, None, None, None, None, 0, None) == 1
