"""CPU model of the blank-speculation schedule of csrc/decoder_ws.cu (the layer-0 epilogue's control update), checked against the
sequential loop of src/asr/decoder_optimized.rs:54-200 on a toy step function.

The kernel runs a stream's decode steps ahead of their vocabulary results (assuming blank) and repairs the stream's state from a
versioned copy when a result breaks the assumption.  This model restates that control logic line by line (same variables:
res / res_prev, kinds bit mask, OP_STEP / OP_COPY, version = tick % W_V) with a hash as the "LSTM state", and asserts that for
random outcomes, random speculation depths per tick and all limits the emitted tokens, step counts and final states equal the
sequential loop's.  It guards the schedule, not the arithmetic (the GPU parity tests do that)."""
import random

W_DMAX, W_V, W_R = 3, 5, 8
BLANK = 1024


def step_fn(state, token, frame, seed):
    """Toy decoder step: new state and the argmax, deterministic in its inputs."""
    h = hash((state, token, frame, seed)) & 0xFFFFFFFF
    r = h % 100
    if r < 70:
        k = BLANK
    elif r < 98:
        k = h % 1024
    else:
        k = 1025 + h % 5  # outside the embedding table: the next step fails
    return h, k


def sequential(length, seed, max_sym, max_total, state0=0):
    """decoder_optimized.rs:54-200 in single-step mode."""
    tokens, state, last, nsteps, failed = [], state0, BLANK, 0, False
    t = 0
    total = 0
    while t < length and total < max_total and not failed:
        sym = 0
        while True:
            sym += 1
            if sym > max_sym:
                break
            if last >= 1025:
                failed = True
                break
            state, k = step_fn(state, last, t, seed)
            nsteps += 1
            if k == BLANK:
                break
            tokens.append(k)
            last = k
            total += 1
            if total >= max_total:
                break
        t += 1
    return tokens, nsteps, state, failed


def speculative(length, seed, max_sym, max_total, depth_of_tick, state0=0):
    """The kernel's schedule for one stream: tick `it` consumes results up to res(it) = max(res(it-1), it-1-d(it))."""
    ver = [None] * W_V
    ver[W_V - 1] = state0
    key = [None] * W_R
    c = dict(t=0, sym=0, total=0, last=BLANK, nsteps=0, tuse=0)
    active, failed, kinds = (1 if length > 0 else 0), 0, 0
    tokens, final = [], (state0 if length == 0 else None)
    res_prev = -1
    it = 0
    while active:
        res = max(res_prev, it - 1 - depth_of_tick(it)) if it > 0 else -1
        assert res <= it - 1 and res >= it - 1 - W_DMAX
        vcur, vprev = it % W_V, (it + W_V - 1) % W_V
        op, src, fin_ver, restore = 0, vprev, -1, False
        for j in range(res_prev + 1, res + 1):
            if not (kinds >> (j & 7)) & 1:
                continue
            kinds &= ~(1 << (j & 7))
            bi = key[j & (W_R - 1)]
            c["nsteps"] += 1
            c["sym"] += 1
            if bi == BLANK:
                c["t"] += 1
                c["sym"] = 0
                if c["t"] >= length:
                    active = 0
            else:
                tokens.append(bi)
                c["total"] += 1
                c["last"] = bi
                if c["total"] >= max_total:
                    active = 0
                elif c["sym"] >= max_sym:
                    c["t"] += 1
                    c["sym"] = 0
                    if c["t"] >= length:
                        active = 0
                if active and bi >= 1025:
                    active, failed = 0, 1
            if not (bi == BLANK and active):
                kinds = 0
                if not active:
                    fin_ver = j % W_V
                elif j < it - 1:
                    restore, src = True, j % W_V
                break
        if active:
            if restore:
                op = 2
            else:
                tspec = c["t"] + bin(kinds).count("1")
                if tspec < length:
                    op, c["tuse"] = 1, tspec
                    kinds |= 1 << (it & 7)
                else:
                    op = 2
        if fin_ver >= 0:
            final = ver[fin_ver]
        if op == 1:
            ver[vcur], key[it & (W_R - 1)] = step_fn(ver[vprev], c["last"], c["tuse"], seed)
        elif op == 2:
            ver[vcur] = ver[src]
            key[it & (W_R - 1)] = None
        res_prev = res
        it += 1
        assert it < 100000
    return tokens, c["nsteps"], final, bool(failed), it


def test_speculative_schedule_equals_sequential_loop():
    rng = random.Random(7)
    ticks = steps = 0
    for trial in range(3000):
        length = rng.choice([0, 1, 2, 3, 7, 20, 60])
        max_sym = rng.choice([1, 2, 30])
        max_total = rng.choice([1, 5, 200])
        seed = rng.randrange(1 << 30)
        mode = rng.choice(["fixed0", "fixed3", "fixed1", "random", "switch"])
        sw = rng.randrange(1, 40)

        def depth(it, mode=mode, sw=sw, r=random.Random(seed)):
            if mode == "fixed0":
                return 0
            if mode == "fixed3":
                return 3
            if mode == "fixed1":
                return 1
            if mode == "switch":
                return 0 if it < sw else W_DMAX
            return r.randrange(0, W_DMAX + 1)

        want = sequential(length, seed, max_sym, max_total, state0=trial)
        got = speculative(length, seed, max_sym, max_total, depth, state0=trial)
        assert got[0] == want[0] and got[1] == want[1] and got[3] == want[3], (trial, length, max_sym, max_total, mode)
        assert got[2] == want[2], (trial, "final state")
        ticks += got[4]
        steps += want[1]
    assert steps > 10000


def test_depth_zero_is_one_tick_per_step():
    for seed in range(50):
        want = sequential(40, seed, 30, 200)
        got = speculative(40, seed, 30, 200, lambda it: 0)
        assert got[:2] == want[:2] and got[4] == want[1] + 1  # one tick per step + the tick that observes the end
