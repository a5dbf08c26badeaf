/* synth_encoder.c — stand-in for the encoder model (out of scope) in bench.py's stream-group leg: an
 * `amira_encoder_fn` (include/amira_b200.h) that returns synthetic encoder outputs [1024][T] of the length the reference's
 * encoder would produce (three stride-2 stages: L <- (L - 1) / 2 + 1), as a window into a table of N(0, 0.5) values made once.
 * O(1) per call, so the leg times the library's stream group, not the stub.  Benchmark infrastructure: not part of libamira_b200. */
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    float *table;
    int64_t n;      /* floats in the table */
    int64_t cursor;
    int64_t calls;
} synth_encoder;

static uint64_t lcg(uint64_t *s) {
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return *s >> 33;
}

synth_encoder *synth_encoder_create(uint64_t seed) {
    synth_encoder *e = (synth_encoder *)calloc(1, sizeof(*e));
    if (!e) return NULL;
    e->n = (int64_t)1 << 22;
    e->table = (float *)malloc(sizeof(float) * (size_t)e->n);
    if (!e->table) { free(e); return NULL; }
    uint64_t s = seed ? seed : 1;
    for (int64_t i = 0; i < e->n; ++i) {  /* sum of four uniforms: variance 1/3 -> scaled to a standard deviation of 0.5 */
        float u = 0.f;
        for (int k = 0; k < 4; ++k) u += (float)(lcg(&s) & 0xFFFFFF) / 16777216.0f;
        e->table[i] = (u - 2.0f) * 0.8660254f;
    }
    return e;
}

void synth_encoder_destroy(synth_encoder *e) {
    if (e) { free(e->table); free(e); }
}

int64_t synth_encoder_calls(const synth_encoder *e) { return e ? e->calls : 0; }

int32_t synth_encoder_fn(void *user, const float *features, int64_t features_len, const float **encoder_outputs, int64_t *encoded_len) {
    synth_encoder *e = (synth_encoder *)user;
    if (!e || !encoder_outputs || !encoded_len || features_len < 0 || (features_len > 0 && !features)) return 1;
    int64_t L = features_len;
    if (L > 0)
        for (int k = 0; k < 3; ++k) L = (L - 1) / 2 + 1;
    const int64_t need = 1024 * L;
    if (need > e->n) return 1;
    if (e->cursor + need > e->n) e->cursor = 0;
    *encoder_outputs = e->table + e->cursor;
    *encoded_len = L;
    e->cursor += need ? need + 64 : 0;
    e->calls += 1;
    return 0;
}
