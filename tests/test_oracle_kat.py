"""Pin the oracle against the reference's own known-answer tests.

Each test restates one reference test: the literal inputs / frozen sums are the
reference's; the brick-tree initialisation order (one shared RandomState walked
children-first, Linear draws bias then W, B/bricks/interfaces.py:157-200) is
emulated by hand and documented next to every draw.
"""
import itertools

import numpy as np
from numpy.testing import assert_allclose

from oracle import lvsr_oracle as O


# ---- tests/test_conv1d.py:6-13 ------------------------------------------------

def test_conv1d_reference_vectors():
    a = np.array([[1.0, 2, 3], [1, 0, 1]])
    b = np.array([[2, 1], [1, 3.0]])
    assert_allclose(O.conv1d(a, b), [[[5, 8], [5, 9]], [[1, 2], [3, 1]]])
    assert_allclose(O.conv1d(a, b, border_mode="full"),
                    [[[2, 5, 8, 3], [1, 5, 9, 9]], [[2, 1, 2, 1], [1, 3, 1, 3]]])


# ---- libs/blocks/tests/bricks/test_recurrent.py:432-453 -----------------------

def test_gru_one_step_closed_form():
    h0 = 0.1 * np.array([[1, 1, 0], [0, 1, 1]], dtype=float)
    x = 0.1 * np.array([[1, 2, 3], [4, 5, 6]], dtype=float)
    zi = (h0 + x) / 2
    ri = -x
    W = 2 * np.ones((3, 3))
    # Constant(2) weights, Tanh gate activation in the reference test
    got = O.gru_step(h0, x, np.hstack([zi, ri]), W, np.hstack([W, W]),
                     activation=np.tanh, gate_activation=np.tanh)
    z = np.tanh(h0.dot(W) + zi)
    r = np.tanh(h0.dot(W) + ri)
    want = z * np.tanh((r * h0).dot(W) + x) + (1 - z) * h0
    assert_allclose(got, want, rtol=1e-6)


# ---- libs/blocks/tests/bricks/test_recurrent.py:455-495 -----------------------

def test_gru_many_steps_masked():
    rng = np.random.RandomState(1)          # GatedRecurrent(..., seed=1), IsotropicGaussian()
    W = rng.normal(0, 1, (3, 3))            # state_to_state
    Wz = rng.normal(0, 1, (3, 3))           # state_to_update
    Wr = rng.normal(0, 1, (3, 3))           # state_to_reset
    x = 0.1 * np.asarray(list(itertools.permutations(range(4))), dtype=float)
    x = np.ones((24, 4, 3)) * x[..., None]
    ri = 0.3 - x
    zi = 2 * ri
    mask = np.ones((24, 4))
    mask[12:24, 3] = 0
    p = dict(state_to_state=W, state_to_gates=np.hstack([Wz, Wr]), initial_state=np.zeros(3))
    got = O.gru_scan(x, np.concatenate([zi, ri], axis=2), mask, p,
                     activation=np.tanh, gate_activation=np.tanh)
    h = np.zeros((25, 4, 3))
    for i in range(1, 25):
        z = np.tanh(h[i - 1].dot(Wz) + zi[i - 1])
        r = np.tanh(h[i - 1].dot(Wr) + ri[i - 1])
        h[i] = np.tanh((r * h[i - 1]).dot(W) + x[i - 1])
        h[i] = z * h[i] + (1 - z) * h[i - 1]
        h[i] = mask[i - 1, :, None] * h[i] + (1 - mask[i - 1, :, None]) * h[i - 1]
    assert_allclose(got, h[1:], rtol=1e-6)
    # masked column keeps its state
    assert_allclose(got[12:, 3], np.repeat(got[11:12, 3], 12, 0))


# ---- libs/blocks/tests/bricks/test_recurrent.py:519-534 (concat order) --------

def test_bidirectional_is_forward_plus_reversed():
    cfg = O.make_config(num_features=5, dims_bidir=[3], dim_dec=4, conv_n=2,
                        conv_num_filters=2, num_phonemes=6)
    params = O.init_params(cfg, seed=3, weights_std=0.5)
    rng = np.random.RandomState(0)
    x = rng.normal(size=(24, 4, 5))
    mask = np.ones((24, 4))
    mask[12:, 3] = 0
    base = "/recognizer/encoder/bidir0"
    y = O.bidirectional(x, mask, params, base)
    fwd = O.recurrent_with_fork(x, mask, params, base + "/forward", reverse=False)
    # backward net applied as a FORWARD net to the reversed input, then re-reversed
    bwd = O.recurrent_with_fork(x[::-1], mask[::-1], params, base + "/backward", reverse=False)
    assert_allclose(y[..., :3], fwd, rtol=1e-12)
    assert_allclose(y[::-1, ..., 3:], bwd, rtol=1e-12)


# ---- libs/blocks/tests/bricks/test_attention.py:61-135 ------------------------

def _rand(rng, size):
    return rng.uniform(size=size)


def _generate_mask(rng, length, batch_size):
    mask = np.ones((length, batch_size))
    for i in range(batch_size):
        mask[1 + rng.randint(0, length - 1):, i] = 0.0
    return mask


def test_attention_recurrent_freeze_sums():
    dim, batch, in_len, att_dim, att_len = 5, 4, 20, 10, 15
    init = np.random.RandomState(1234)       # AttentionRecurrent(..., seed=1234)
    g = lambda shape: init.normal(0, 0.5, size=shape)   # IsotropicGaussian(0.5)
    # children = [transition, attention, distribute] (B/bricks/attention.py:577)
    W_rec = g((dim, dim))                    # SimpleRecurrent.W
    W_state = g((dim, att_dim))              # state_trans/transform_states.W (match_dim = attended_dim)
    W_pre = g((att_dim, att_dim))            # preprocess: bias (Constant, no draw) then W
    v = g((att_dim, 1))                      # energy_comp/linear.W
    W_dist = g((att_dim, dim))               # distribute/fork_inputs.W

    rng = np.random.RandomState(1234)
    inputs = _rand(rng, (in_len, batch, dim))
    inputs_mask = _generate_mask(rng, in_len, batch)
    attended = _rand(rng, (att_len, batch, att_dim))
    attended_mask = _generate_mask(rng, att_len, batch)

    P = attended.dot(W_pre)
    s = np.zeros((batch, dim))
    states, glimpses, weights = [], [], []
    for t in range(in_len):
        # take_glimpses from the PREVIOUS state, then distribute, then transition
        ctx, w = O.content_take_glimpses(attended, P, attended_mask, s, W_state, v)
        s = O.simple_recurrent_step(s, inputs[t] + ctx.dot(W_dist), W_rec, inputs_mask[t],
                                    activation=lambda z: z)   # Identity activation
        states.append(s); glimpses.append(ctx); weights.append(w)
    states, glimpses, weights = map(np.stack, (states, glimpses, weights))

    assert np.all(weights * (1 - attended_mask.T) == 0)
    assert np.all(abs(weights + (1 - attended_mask.T)) > 1e-5)
    for i in range(batch):
        last = int(inputs_mask[:, i].sum())
        for j in range(last, in_len):
            assert_allclose(weights[last, i], weights[j, i], 1e-5)
    assert_allclose(weights.sum(), in_len * batch, 1e-5)
    assert_allclose(states.sum(), 113.429, rtol=1e-5)
    assert_allclose(glimpses.sum(), 415.901, rtol=1e-5)


# ---- libs/blocks/tests/bricks/test_attention.py:138-182 -----------------------

def test_compute_weights_invariants():
    rng = np.random.RandomState(0)
    e = rng.rand(5, 6)
    assert np.all(np.isfinite(O.compute_weights(e, np.zeros((5, 6)))))
    big = 50.0 * rng.randn(5, 6) + 800
    w = O.compute_weights(big, np.ones((5, 6)))
    assert np.all(np.isfinite(w))
    assert_allclose(w.sum(axis=0), 1.0)


# ---- libs/blocks/tests/bricks/test_sequence_generators.py:96-171 --------------

def test_integer_sequence_generator_freeze_sums():
    readout_dim, feedback_dim, dim, batch, n_steps = 5, 3, 20, 30, 10
    init = np.random.RandomState(1234)       # SequenceGenerator(..., seed=1234)
    g = lambda shape: init.normal(0, 0.1, size=shape)   # IsotropicGaussian(0.1) pushed to ALL children
    # generator.children = [readout, fork, transition]; readout.children =
    # [emitter, feedback_brick, merge, post_merge]
    lookup = g((readout_dim, feedback_dim))  # lookupfeedback/lookuptable.W
    W_merge = g((dim, readout_dim))          # merge/transform_states.W
    b_post = np.zeros(readout_dim)           # post_merge Bias, Constant(0)
    W_fi = g((feedback_dim, dim))            # fork_inputs: b (no draw), W
    W_fg = g((feedback_dim, 2 * dim))        # fork_gate_inputs
    # The Orthogonal() given to the GRU is overwritten by the generator's push
    W_ss = g((dim, dim))                     # state_to_state (recurrent_weights_init falls back to weights_init)
    W_su = g((dim, dim))                     # state_to_update
    W_sr = g((dim, dim))                     # state_to_reset
    gru = dict(state_to_state=W_ss, state_to_gates=np.hstack([W_su, W_sr]),
               initial_state=np.zeros(dim))

    rng = np.random.RandomState(1234)
    y = rng.randint(readout_dim, size=(n_steps, batch))
    mask = np.ones((n_steps, batch))

    def costs_fun(y, mask):
        y = np.asarray(y); mask = np.asarray(mask, dtype=float)
        fb = lookup[y]
        states = O.gru_scan(fb.dot(W_fi), fb.dot(W_fg), mask, gru)
        prev = np.concatenate([np.zeros((1,) + states.shape[1:]), states[:-1]])   # states[:-1] incl. initial
        logp = O.log_softmax(prev.dot(W_merge) + b_post)
        return -np.take_along_axis(logp, y[..., None], -1)[..., 0] * mask

    costs = costs_fun(y, mask)
    assert costs.shape == (n_steps, batch)
    assert_allclose(costs.sum(), 482.827, rtol=1e-5)
    assert_allclose(costs.sum(axis=0).mean(), 16.0942, rtol=1e-5)
    assert_allclose(costs.sum() / mask.sum(), 1.60942, rtol=1e-5)
    # mask-agnostic cost (:167-171)
    c1 = costs_fun([[1], [2]], [[1], [1]])
    c2 = costs_fun([[3, 1], [4, 2], [2, 0]], [[1, 1], [1, 1], [1, 0]])
    assert_allclose(c1.sum(), c2[:, 1].sum(), rtol=1e-5)


# ---- libs/blocks/tests/bricks/test_sequence_generators.py:197-275 -------------

def test_sequence_generator_with_attention_freeze_sums():
    inp_dim, inp_len, att_dim, att_len, batch, n_steps = 2, 10, 3, 11, 4, 30
    rng = np.random.RandomState(1234)
    outputs = _rand(rng, (inp_len, batch, inp_dim))
    outputs_mask = _generate_mask(rng, inp_len, batch)
    attended = _rand(rng, (att_len, batch, att_dim))
    attended_mask = _generate_mask(rng, att_len, batch)

    init = np.random.RandomState(1234)
    g = lambda shape: init.normal(0, 0.1, size=shape)
    # readout: emitter, feedback (trivial), merge [states, weighted_averages], post_merge Bias
    W_ms = g((inp_dim, inp_dim))             # merge/transform_states.W
    W_mw = g((att_dim, inp_dim))             # merge/transform_weighted_averages.W
    b_post = np.zeros(inp_dim)
    W_fork = g((inp_dim, inp_dim))           # fork/fork_inputs: b (no draw), W
    # att_trans.children = [transition, attention, distribute]
    W_rec = g((inp_dim, inp_dim))            # TestTransition (SimpleRecurrent).W
    W_state = g((inp_dim, inp_dim))          # state_trans/transform_states.W  (match_dim = inp_dim)
    W_pre = g((att_dim, inp_dim))            # preprocess.W
    v = g((inp_dim, 1))                      # energy_comp/linear.W
    W_dist = g((att_dim, inp_dim))           # distribute/fork_inputs.W

    ident = lambda z: z
    P = attended.dot(W_pre)

    # cost_matrix: teacher forcing, feedback = the outputs themselves
    inputs = outputs.dot(W_fork)
    s = np.zeros((batch, inp_dim))
    prev_states, ctxs = [], []
    for i in range(inp_len):
        prev_states.append(s)
        ctx, w = O.content_take_glimpses(attended, P, attended_mask, s, W_state, v)
        s = O.simple_recurrent_step(s, inputs[i] + ctx.dot(W_dist), W_rec, outputs_mask[i], ident)
        ctxs.append(ctx)
    readouts = np.stack(prev_states).dot(W_ms) + np.stack(ctxs).dot(W_mw) + b_post
    costs = ((readouts - outputs) ** 2).sum(axis=-1) * outputs_mask
    assert_allclose(costs.sum(), 13.5042, rtol=1e-5)

    # generate: y_0 = 0, s_0 = 0; glimpses -> readout -> emit (identity) -> fork -> next state
    s = np.zeros((batch, inp_dim))
    S, Y, G, Wts = [], [], [], []
    for i in range(n_steps):
        ctx, w = O.content_take_glimpses(attended, P, attended_mask, s, W_state, v)
        y = s.dot(W_ms) + ctx.dot(W_mw) + b_post
        s = O.simple_recurrent_step(s, y.dot(W_fork) + ctx.dot(W_dist), W_rec, None, ident)
        S.append(s); Y.append(y); G.append(ctx); Wts.append(w)
    assert_allclose(np.sum(S), 23.4172, rtol=1e-5)
    assert_allclose(np.sum(Wts), 120.0, rtol=1e-5)
    assert_allclose(np.sum(G), 199.2402, rtol=1e-5)
    assert_allclose(np.sum(Y), -11.6008, rtol=1e-5)


# ---- libs/blocks/tests/test_search.py:65-69 -----------------------------------

def test_beam_search_smallest():
    a = np.array([[3, 6, 4], [1, 2, 7]])
    ind, mins = O.smallest(a, 2)
    assert np.all(np.array(ind) == np.array([[1, 1], [0, 1]]))
    assert np.all(mins == [1, 2])


# ---- libs/blocks/tests/bricks/test_bricks.py (Maxout literal semantics) --------

def test_maxout_adjacent_pieces():
    x = np.arange(12, dtype=float).reshape(2, 6)
    assert_allclose(O.maxout(x, 2), [[1, 3, 5], [7, 9, 11]])
    assert_allclose(O.maxout(-x, 3), [[0, -3], [-6, -9]])


def test_log_softmax_literal():
    x = np.array([[1.0, 2.0, 3.0]])
    want = x - np.log(np.exp(x).sum())
    assert_allclose(O.log_softmax(x), want, rtol=1e-12)
