// host_common.h — types shared by the host-side translation units above the C ABI (host_pipeline.cpp, host_stream.cpp):
// Vocabulary (src/asr/types.rs:77-155) and the pipeline object behind `amira_pipeline`.
#pragma once
#include <cerrno>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "amira_b200.h"

namespace amira_host {

inline bool is_space(unsigned char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\v' || ch == '\f'; }

// Vocabulary::load_from_file (src/asr/types.rs:87-108): "<token> <id>" per line; token = all fields but the last.
struct Vocabulary {
    std::unordered_map<int32_t, std::string> id_to_token;

    bool load(const std::string &path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) return false;
        std::string line;
        while (std::getline(f, line)) {
            std::vector<std::string> parts;
            size_t i = 0;
            while (i < line.size()) {
                while (i < line.size() && is_space((unsigned char)line[i])) ++i;
                size_t j = i;
                while (j < line.size() && !is_space((unsigned char)line[j])) ++j;
                if (j > i) parts.emplace_back(line.substr(i, j - i));
                i = j;
            }
            if (parts.size() < 2) continue;
            const std::string &id_s = parts.back();
            char *end = nullptr;
            errno = 0;
            const long long v = std::strtoll(id_s.c_str(), &end, 10);
            if (errno || end == id_s.c_str() || *end != '\0' || v < INT32_MIN || v > INT32_MAX) continue;
            if (id_s[0] == '+' && id_s.size() == 1) continue;
            std::string tok = parts[0];
            for (size_t k = 1; k + 1 < parts.size(); ++k) tok += " " + parts[k];
            id_to_token[(int32_t)v] = tok;  // later lines overwrite (HashMap::insert)
        }
        return true;
    }

    // Vocabulary::decode_tokens (src/asr/types.rs:111-135)
    std::string decode(const int32_t *ids, int32_t n) const {
        static const char kSp[] = "\xE2\x96\x81";  // U+2581
        std::string out;
        for (int32_t i = 0; i < n; ++i) {
            auto it = id_to_token.find(ids[i]);
            if (it == id_to_token.end()) continue;  // unknown ids are skipped silently (:115-116)
            const std::string &t = it->second;
            if (t.compare(0, 3, kSp) == 0) {
                out += ' ';
                out.append(t, 3, std::string::npos);
            } else {
                out += t;
            }
        }
        size_t a = 0, b = out.size();
        while (a < b && is_space((unsigned char)out[a])) ++a;
        while (b > a && is_space((unsigned char)out[b - 1])) --b;
        return out.substr(a, b - a);
    }
};

}  // namespace amira_host

struct amira_pipeline {
    amira_ctx *ctx = nullptr;
    amira_encoder_fn encoder = nullptr;
    void *encoder_user = nullptr;
    amira_host::Vocabulary vocab;
    std::mutex mu;
    std::string err;
    std::vector<float> features, wave;
    std::vector<int32_t> tokens;
};
