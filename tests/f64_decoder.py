"""Float64 numpy restatement of the greedy loop (src/asr/decoder_optimized.rs:54-200) around the LSTM / joint step — TEST
INFRASTRUCTURE.  It exists to CALIBRATE state comparisons: over ~500 sequential steps the recurrent state of any fp32
implementation drifts from exact arithmetic by an amount set by the model's own sensitivity, so the GPU's distance from the
fp32 oracle is judged against the fp32 oracle's own distance from this float64 run (same tokens, same control flow)."""
import numpy as np

H, ENC, V, BLANK = 640, 1024, 1030, 1024


def _views(blob):
    o, t = 0, {}

    def take(name, *shape):
        nonlocal o
        n = int(np.prod(shape))
        t[name] = blob[o:o + n].reshape(shape).astype(np.float64)
        o += n

    take("emb", 1025, H)
    for l in range(2):
        take(f"w_ih{l}", 4 * H, H); take(f"w_hh{l}", 4 * H, H); take(f"b_ih{l}", 4 * H); take(f"b_hh{l}", 4 * H)
    take("w_enc", H, ENC); take("b_enc", H); take("w_pred", H, H); take("b_pred", H); take("w_out", V, H); take("b_out", V)
    return t


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def greedy_decode_f64(blob, enc, lens, max_symbols=30, max_total=200):
    """enc [B,1024,T] f32, lens [B].  Returns tokens (list of lists), n_steps [B], states_1/2 [2,B,640] float64, and the smallest
    top-1/top-2 logit margin seen per stream."""
    w = _views(np.asarray(blob, np.float32))
    enc = np.asarray(enc, np.float64)
    B, _, T = enc.shape
    E = np.einsum("hf,bft->bth", w["w_enc"], enc) + w["b_enc"] + w["b_pred"]
    h = np.zeros((2, B, H)); c = np.zeros((2, B, H))
    t = np.zeros(B, np.int64); sym = np.zeros(B, np.int64); last = np.full(B, BLANK, np.int64)
    lens = np.asarray(lens, np.int64)
    active = lens > 0
    toks = [[] for _ in range(B)]
    nsteps = np.zeros(B, np.int64)
    margin = np.full(B, np.inf)
    while active.any():
        idx = np.nonzero(active)[0]
        x = w["emb"][last[idx]]
        for l in range(2):
            g = x @ w[f"w_ih{l}"].T + w[f"b_ih{l}"] + h[l, idx] @ w[f"w_hh{l}"].T + w[f"b_hh{l}"]
            i, f, gg, o = _sig(g[:, :H]), _sig(g[:, H:2 * H]), np.tanh(g[:, 2 * H:3 * H]), _sig(g[:, 3 * H:])
            cn = f * c[l, idx] + i * gg
            hn = o * np.tanh(cn)
            c[l, idx] = cn; h[l, idx] = hn
            x = hn
        z = np.tanh(E[idx, t[idx]] + x @ w["w_pred"].T)
        logits = z @ w["w_out"].T + w["b_out"]
        k = logits.argmax(axis=1)
        srt = np.sort(logits, axis=1)
        margin[idx] = np.minimum(margin[idx], srt[:, -1] - srt[:, -2])
        nsteps[idx] += 1
        for j, b in enumerate(idx):
            sym[b] += 1
            if k[j] == BLANK:
                t[b] += 1; sym[b] = 0
                if t[b] >= lens[b]: active[b] = False
            else:
                toks[b].append(int(k[j])); last[b] = k[j]
                if len(toks[b]) >= max_total: active[b] = False
                elif sym[b] >= max_symbols:
                    t[b] += 1; sym[b] = 0
                    if t[b] >= lens[b]: active[b] = False
    return toks, nsteps, h, c, margin
