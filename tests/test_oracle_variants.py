"""NON-REFERENCE decoding variants of the oracle (SURVEY.md 8c "optional flags", 8 f4): the canonical RNN-T state rule and the TDT
reading of the 1030 outputs.  CPU only; they document what the flags mean and pin them with mock steps, so that a later GPU
implementation has an oracle to be checked against.  The product library implements the reference's literal loop only."""
import numpy as np

BLANK = 1024


def _mock(seq):
    """A step that replays (token, duration) decisions and counts state updates in states_1[0]."""
    it = iter(seq)

    def step(frame, targets, s1, s2):
        k, d = next(it)
        lg = np.full(1030, -10.0, np.float32)
        lg[k] = 5.0
        lg[1025 + d] = 7.0   # duration logit: larger than every token logit on purpose (a flat 1030-way argmax would pick it)
        s1 = s1.copy()
        s1[0] += 1.0
        return lg, s1, s2

    return step


def test_canonical_rule_keeps_the_state_on_blank(oracle):
    enc = np.zeros((1024, 4), np.float32)
    seq = [(5, 0), (BLANK, 0), (BLANK, 0), (7, 0), (BLANK, 0), (BLANK, 0)]
    # reference rule: flat argmax over all 1030 outputs would see the duration logit — use plain logits for this part
    def plain(seq):
        it = iter(seq)
        def step(frame, targets, s1, s2):
            k, _ = next(it)
            lg = np.full(1030, -10.0, np.float32); lg[k] = 5.0
            s1 = s1.copy(); s1[0] += 1.0
            return lg, s1, s2
        return step
    ref = oracle.greedy_decode(enc, 4, step=plain(seq))
    assert ref.tokens == [5, 7] and ref.n_steps == 6 and ref.states_1.ravel()[0] == 6.0       # state replaced after EVERY step (:154)
    can = oracle.greedy_decode(enc, 4, step=plain(seq), state_update_on_nonblank_only=True)
    assert can.tokens == [5, 7] and can.n_steps == 6 and can.states_1.ravel()[0] == 2.0       # only the two emitting steps count


def test_tdt_durations_move_the_frame_index(oracle):
    enc = np.zeros((1024, 10), np.float32)
    # t=0: token 5 with duration 2 -> t=2: token 6, duration 0 (stay) -> blank, duration 0 (forced to 1) -> t=3: token 7, duration 4 -> t=7:
    # blank duration 3 -> t=10: end
    seq = [(5, 2), (6, 0), (BLANK, 0), (7, 4), (BLANK, 3)]
    r = oracle.greedy_decode(enc, 10, step=_mock(seq), tdt_durations=True)
    assert r.tokens == [5, 6, 7] and r.n_steps == 5 and r.frames_visited == 4
    # the reference's flat 1030-way argmax on the same logits picks the duration outputs instead: index 1027 is pushed as a token
    flat = oracle.greedy_decode(enc, 10, step=_mock(seq))
    assert flat.tokens[0] == 1027


def test_tdt_with_the_model_is_a_valid_decode(oracle):
    m = oracle.Model(seed=3456)
    enc = (0.5 * np.random.default_rng(1).standard_normal((1024, 30))).astype(np.float32)
    r = oracle.greedy_decode(enc, 30, m, tdt_durations=True, state_update_on_nonblank_only=True)
    assert r.rc == 0 and all(0 <= t < BLANK for t in r.tokens) and r.frames_visited <= 30


def test_initial_last_token_resumes_a_call(oracle):
    """Two calls with the LSTM state and the last token carried == one call over all frames (what amira_greedy_decode_resume does)."""
    import amira_b200 as A
    m = oracle.Model(blob=A.synthetic_weights(3456))
    enc = np.ascontiguousarray((0.5 * np.random.default_rng(2345).standard_normal((2, 1024, 40))).astype(np.float32)[1])
    one = oracle.greedy_decode(enc, 40, m)
    a = oracle.greedy_decode(np.ascontiguousarray(enc[:, :17]), 17, m)
    last = a.tokens[-1] if a.tokens else BLANK
    b = oracle.greedy_decode(np.ascontiguousarray(enc[:, 17:]), 23, m, states=(a.states_1, a.states_2), initial_last=last)
    assert len(one.tokens) > 0 and a.tokens + b.tokens == one.tokens
    assert np.array_equal(b.states_1, one.states_1)
