// host_wire.cpp — wire formats either side of the hot path, behind the C ABI:
//   * WebSocket binary frames: audio chunk vs. control byte, with the validation order of StreamProcessor::handle_audio_chunk
//     (src/server/stream.rs:215-281);
//   * the JSON body of POST /v2/decode/batch/{model}: `BatchRequest` + `validate()` (src/server/handlers.rs:44-116);
//   * the JSON the server answers with: `AsrResponse` / `StreamStatus` (src/asr/types.rs:236-272), metadata of the batch handler
//     (src/server/handlers.rs:192-206).
// The reference disagrees with itself about the control bytes: the server matches the constants of src/constants.rs:243-246
// (END 0xFF, KEEPALIVE 0x00; src/server/stream.rs:24-26, 234-254) while src/config.rs:95-98, README.md:279-280 and
// examples/simple_client.rs:87 say END 0x00, KEEPALIVE 0x01.  Both dialects are implemented and the caller picks one; dialect
// AMIRA_WIRE_DIALECT_SERVER is what a running reference server accepts, so it is the drop-in default.  With it, a client that
// follows the documentation ends a stream by sending 0x00 — which the server reads as KEEPALIVE — and its keepalive 0x01 is
// answered with "Unknown control byte"; tests/test_wire.py pins exactly that.
// Pure host code: no CUDA, no allocation beyond std::string, no exception crosses the boundary.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "amira_b200.h"

namespace {

constexpr size_t kMaxChunkBytes = 1024 * 1024;         // src/server/stream.rs:221
constexpr size_t kMaxAudioBytes = 100 * 1024 * 1024;   // src/server/handlers.rs:84
constexpr double kMaxBatchSeconds = 30.0;              // src/constants.rs:24
constexpr size_t kMaxOpaqueBytes = 10000;              // src/server/handlers.rs:109

void set_err(char *err, size_t cap, const std::string &m) {
    if (!err || !cap) return;
    const size_t n = m.size() < cap - 1 ? m.size() : cap - 1;
    std::memcpy(err, m.data(), n);
    err[n] = '\0';
}

// ---- a minimal JSON reader: enough to walk one object, pull an array of small integers and skip everything else ----
struct Reader {
    const char *s;
    size_t n, i = 0;
    std::string error;
    bool fail(const std::string &m) {
        if (error.empty()) error = m + " at byte " + std::to_string(i);
        return false;
    }
    void ws() {
        while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
    }
    bool lit(const char *w) {
        const size_t L = std::strlen(w);
        if (i + L > n || std::memcmp(s + i, w, L) != 0) return fail("invalid literal");
        i += L;
        return true;
    }
    bool string(std::string *out) {  // decodes escapes when `out` is given (keys); \uXXXX is kept verbatim
        if (i >= n || s[i] != '"') return fail("expected string");
        ++i;
        while (i < n) {
            const unsigned char c = (unsigned char)s[i];
            if (c == '"') { ++i; return true; }
            if (c < 0x20) return fail("control character in string");
            if (c == '\\') {
                if (i + 1 >= n) return fail("unterminated escape");
                const char e = s[i + 1];
                if (e == 'u') {
                    if (i + 6 > n) return fail("short \\u escape");
                    for (int k = 2; k < 6; ++k)
                        if (!std::isxdigit((unsigned char)s[i + k])) return fail("bad \\u escape");
                    if (out) out->append(s + i, 6);
                    i += 6;
                    continue;
                }
                const char *from = "\"\\/bfnrt", *to = "\"\\/\b\f\n\r\t";
                const char *p = std::strchr(from, e);
                if (!p || !e) return fail("bad escape");
                if (out) out->push_back(to[p - from]);
                i += 2;
                continue;
            }
            if (out) out->push_back((char)c);
            ++i;
        }
        return fail("unterminated string");
    }
    bool number(double *v, bool *is_int) {
        const size_t b = i;
        if (i < n && s[i] == '-') ++i;
        if (i >= n || !std::isdigit((unsigned char)s[i])) return fail("expected number");
        if (s[i] == '0') ++i;
        else while (i < n && std::isdigit((unsigned char)s[i])) ++i;
        bool integral = true;
        if (i < n && s[i] == '.') {
            integral = false;
            ++i;
            if (i >= n || !std::isdigit((unsigned char)s[i])) return fail("bad fraction");
            while (i < n && std::isdigit((unsigned char)s[i])) ++i;
        }
        if (i < n && (s[i] == 'e' || s[i] == 'E')) {
            integral = false;
            ++i;
            if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
            if (i >= n || !std::isdigit((unsigned char)s[i])) return fail("bad exponent");
            while (i < n && std::isdigit((unsigned char)s[i])) ++i;
        }
        if (v) *v = std::strtod(std::string(s + b, i - b).c_str(), nullptr);
        if (is_int) *is_int = integral;
        return true;
    }
    bool skip(int depth = 0) {  // any value
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (i >= n) return fail("unexpected end");
        const char c = s[i];
        if (c == '"') return string(nullptr);
        if (c == 't') return lit("true");
        if (c == 'f') return lit("false");
        if (c == 'n') return lit("null");
        if (c == '[' || c == '{') {
            const char close = c == '[' ? ']' : '}';
            ++i;
            ws();
            if (i < n && s[i] == close) { ++i; return true; }
            for (;;) {
                if (c == '{') {
                    ws();
                    if (!string(nullptr)) return false;
                    ws();
                    if (i >= n || s[i] != ':') return fail("expected ':'");
                    ++i;
                }
                if (!skip(depth + 1)) return false;
                ws();
                if (i < n && s[i] == ',') { ++i; continue; }
                if (i < n && s[i] == close) { ++i; return true; }
                return fail("expected ',' or closing bracket");
            }
        }
        return number(nullptr, nullptr);
    }
};

void json_escape(std::string &out, const char *s) {  // serde_json's escaping: \" \\ \b \f \n \r \t, \u00XX below 0x20, UTF-8 as is
    for (; *s; ++s) {
        const unsigned char c = (unsigned char)*s;
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    std::snprintf(buf, sizeof(buf), "\\u%04x", c);
                    out += buf;
                } else {
                    out.push_back((char)c);
                }
        }
    }
}

}  // namespace

extern "C" {

// StreamProcessor::handle_audio_chunk (src/server/stream.rs:215-281), up to the point where the chunk is buffered
int32_t amira_wire_classify_frame(const uint8_t *data, size_t n, int32_t dialect, int32_t *kind) {
    if (!kind || (n && !data) || (dialect != AMIRA_WIRE_DIALECT_SERVER && dialect != AMIRA_WIRE_DIALECT_DOCUMENTED)) return AMIRA_ERR_INVALID_VALUE;
    const uint8_t end_byte = dialect == AMIRA_WIRE_DIALECT_SERVER ? 0xFF : 0x00;
    const uint8_t keepalive_byte = dialect == AMIRA_WIRE_DIALECT_SERVER ? 0x00 : 0x01;
    if (n > kMaxChunkBytes) *kind = AMIRA_FRAME_TOO_LARGE;               // :220-228
    else if (n == 1) {                                                  // :234-254
        if (data[0] == end_byte) *kind = AMIRA_FRAME_END;
        else if (data[0] == keepalive_byte) *kind = AMIRA_FRAME_KEEPALIVE;
        else *kind = AMIRA_FRAME_UNKNOWN_CONTROL;
    } else if (n % 2 != 0) *kind = AMIRA_FRAME_ODD_LENGTH;              // :257-261
    else if (n == 0) *kind = AMIRA_FRAME_EMPTY;                         // :264-268
    else *kind = AMIRA_FRAME_AUDIO;
    return AMIRA_OK;
}

// serde's view of `BatchRequest` (src/server/handlers.rs:44-64) + BatchRequest::validate (:68-116).  `audio_buffer` is a JSON
// array of integers 0..255 (Vec<u8>); `opaque` is returned as the raw span of its value inside `json`; unknown keys are ignored.
// Returns AMIRA_ERR_INVALID_VALUE with the reference's validation message in `err` when the request is refused.  When
// `audio_cap` is too small the needed size is left in *n_audio and the call fails with AMIRA_ERR_OUT_OF_MEMORY.
int32_t amira_wire_parse_batch_request(const char *json, size_t n, uint8_t *audio, size_t audio_cap, size_t *n_audio,
                                       size_t *opaque_begin, size_t *opaque_len, char *err, size_t err_cap) {
    if (!json || !n_audio) return AMIRA_ERR_INVALID_VALUE;
    try {
        *n_audio = 0;
        if (opaque_begin) *opaque_begin = 0;
        if (opaque_len) *opaque_len = 0;
        Reader r{json, n};
        auto bad = [&](const std::string &m) {
            set_err(err, err_cap, m);
            return (int32_t)AMIRA_ERR_INVALID_VALUE;
        };
        r.ws();
        if (r.i >= n || json[r.i] != '{') return bad("invalid JSON: expected an object");
        ++r.i;
        bool have_audio = false, have_opaque = false, overflow = false;
        size_t count = 0, o_begin = 0, o_len = 0;
        r.ws();
        if (r.i < n && json[r.i] == '}') {
            ++r.i;
        } else {
            for (;;) {
                r.ws();
                std::string key;
                if (!r.string(&key)) return bad("invalid JSON: " + r.error);
                r.ws();
                if (r.i >= n || json[r.i] != ':') return bad("invalid JSON: expected ':'");
                ++r.i;
                r.ws();
                if (key == "audio_buffer") {
                    if (have_audio) return bad("duplicate field `audio_buffer`");
                    have_audio = true;
                    if (r.i >= n || json[r.i] != '[') return bad("invalid type: audio_buffer must be an array of bytes");
                    ++r.i;
                    r.ws();
                    if (r.i < n && json[r.i] == ']') {
                        ++r.i;
                    } else {
                        for (;;) {
                            r.ws();
                            double v = 0;
                            bool is_int = false;
                            if (!r.number(&v, &is_int)) return bad("invalid type: audio_buffer must hold integers 0..255");
                            if (!is_int || v < 0 || v > 255) return bad("invalid value: audio_buffer must hold integers 0..255");
                            if (audio && count < audio_cap) audio[count] = (uint8_t)v;
                            else overflow = true;
                            ++count;
                            r.ws();
                            if (r.i < n && json[r.i] == ',') { ++r.i; continue; }
                            if (r.i < n && json[r.i] == ']') { ++r.i; break; }
                            return bad("invalid JSON: expected ',' or ']' in audio_buffer");
                        }
                    }
                } else if (key == "opaque") {
                    if (have_opaque) return bad("duplicate field `opaque`");
                    have_opaque = true;
                    o_begin = r.i;
                    if (!r.skip()) return bad("invalid JSON: " + r.error);
                    o_len = r.i - o_begin;
                    if (o_len == 4 && std::memcmp(json + o_begin, "null", 4) == 0) o_len = 0;  // Option::None
                } else if (!r.skip()) {
                    return bad("invalid JSON: " + r.error);
                }
                r.ws();
                if (r.i < n && json[r.i] == ',') { ++r.i; continue; }
                if (r.i < n && json[r.i] == '}') { ++r.i; break; }
                return bad("invalid JSON: expected ',' or '}'");
            }
        }
        r.ws();
        if (r.i != n) return bad("invalid JSON: trailing characters");
        if (!have_audio) return bad("missing field `audio_buffer`");
        *n_audio = count;
        // BatchRequest::validate, in its order
        if (count == 0) return bad("Audio buffer cannot be empty");
        if (count % 2 != 0) return bad("Audio buffer length must be even for 16-bit PCM");
        if (count > kMaxAudioBytes)
            return bad("Audio buffer too large: " + std::to_string(count) + " bytes (max: " + std::to_string(kMaxAudioBytes) + " bytes)");
        const float secs = (float)count / (16000.0f * 2.0f);
        if (secs > (float)kMaxBatchSeconds) {
            char buf[96];
            std::snprintf(buf, sizeof(buf), "Audio too long: %.1fs (max: %ds)", (double)secs, (int)kMaxBatchSeconds);
            return bad(buf);
        }
        if (o_len > kMaxOpaqueBytes) return bad("Opaque data too large (max: 10KB)");
        if (opaque_begin) *opaque_begin = o_begin;
        if (opaque_len) *opaque_len = o_len;
        if (overflow || (!audio && count)) {
            set_err(err, err_cap, "audio buffer too small");
            return AMIRA_ERR_OUT_OF_MEMORY;
        }
        return AMIRA_OK;
    } catch (...) {
        set_err(err, err_cap, "host allocation failed");
        return AMIRA_ERR_OUT_OF_MEMORY;
    }
}

// `AsrResponse` (src/asr/types.rs:251-272, camelCase, None fields omitted) with the batch handler's metadata
// (src/server/handlers.rs:192-206) when `meta` is given.  status: 0 ACTIVE, 1 COMPLETE, 2 PAUSED, 3 ERROR (UPPERCASE on the wire).
int32_t amira_wire_format_response(const char *transcription, int32_t status, const char *message, const amira_transcription *meta,
                                   const int32_t *tokens, const char *opaque_json, char *out, size_t out_cap, size_t *out_len) {
    if (!transcription || status < 0 || status > 3 || (meta && meta->n_tokens > 0 && !tokens)) return AMIRA_ERR_INVALID_VALUE;
    try {
        static const char *kStatus[4] = {"ACTIVE", "COMPLETE", "PAUSED", "ERROR"};
        std::string s = "{\"transcription\":\"";
        json_escape(s, transcription);
        s += "\",\"status\":\"";
        s += kStatus[status];
        s += "\"";
        if (message) {
            s += ",\"message\":\"";
            json_escape(s, message);
            s += "\"";
        }
        if (meta) {
            s += ",\"metadata\":{\"audio_length_samples\":" + std::to_string(meta->audio_length_samples) +
                 ",\"features_length\":" + std::to_string(meta->features_length) + ",\"encoded_length\":" + std::to_string(meta->encoded_length) +
                 ",\"tokens\":[";
            for (int32_t i = 0; i < meta->n_tokens; ++i) {
                if (i) s += ",";
                s += std::to_string(tokens[i]);
            }
            s += "]}";
        }
        if (opaque_json && *opaque_json) {
            s += ",\"opaque\":";
            s += opaque_json;
        }
        s += "}";
        if (out_len) *out_len = s.size();
        if (out && out_cap) {
            const size_t m = s.size() < out_cap - 1 ? s.size() : out_cap - 1;
            std::memcpy(out, s.data(), m);
            out[m] = '\0';
            if (m < s.size()) return AMIRA_ERR_OUT_OF_MEMORY;
        }
        return AMIRA_OK;
    } catch (...) {
        return AMIRA_ERR_OUT_OF_MEMORY;
    }
}

}  // extern "C"
