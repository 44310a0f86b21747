"""Pin the oracle's COMPOSITE functions -- the ones the CUDA path is compared with
(O.cost_matrix, O.take_glimpses, O.compute_energies, O.compute_states, O.readout,
O.encoder) -- to the reference's frozen sums, by driving the reference's known-answer
tests through them with degenerate parameters (a zero handler makes the location term
vanish, a zero distribute / transform_weighted_averages removes the attention from a
generator that has none, linear_0.W = I with the identity activation turns the
post-merge MLP into the bare Bias of the reference test).

Same literals and brick-initialisation order as tests/test_oracle_kat.py.
"""
import itertools
from collections import OrderedDict

import numpy as np
from numpy.testing import assert_allclose

from oracle import lvsr_oracle as O

ATT = "/recognizer/generator/att_trans/conv_att"
TR = "/recognizer/generator/att_trans"
GEN = "/recognizer/generator"


def _rand(rng, size):
    return rng.uniform(size=size)


def _generate_mask(rng, length, batch_size):
    mask = np.ones((length, batch_size))
    for i in range(batch_size):
        mask[1 + rng.randint(0, length - 1):, i] = 0.0
    return mask


def _attention_params(W_state, W_pre, v, K, n, seed=99):
    """conv attention that degenerates to SequenceContentAttention: handler.W = 0
    (lvsr/bricks/attention.py:108-111 adds conv . handler.W to the match vector)."""
    rng = np.random.RandomState(seed)
    M = W_state.shape[1]
    return {
        ATT + "/state_trans/transform_states.W": W_state,
        ATT + "/preprocess.W": W_pre,
        ATT + "/preprocess.b": np.zeros(M),
        ATT + "/energy_comp/linear.W": v,
        ATT + "/handler.W": np.zeros((K, M)),
        ATT + "/conv1d.filters": rng.normal(size=(K, 2 * n + 1)),   # arbitrary: multiplied by zero
    }


# ---- libs/blocks/tests/bricks/test_attention.py:61-135 through O.take_glimpses ------------

def test_attention_freeze_sums_through_take_glimpses():
    dim, batch, in_len, att_dim, att_len = 5, 4, 20, 10, 15
    init = np.random.RandomState(1234)
    g = lambda shape: init.normal(0, 0.5, size=shape)
    W_rec = g((dim, dim))
    W_state = g((dim, att_dim))
    W_pre = g((att_dim, att_dim))
    v = g((att_dim, 1))
    W_dist = g((att_dim, dim))

    rng = np.random.RandomState(1234)
    inputs = _rand(rng, (in_len, batch, dim))
    inputs_mask = _generate_mask(rng, in_len, batch)
    attended = _rand(rng, (att_len, batch, att_dim))
    attended_mask = _generate_mask(rng, att_len, batch)

    cfg = O.make_config(num_features=3, dims_bidir=[att_dim // 2], dim_dec=dim, dim_matcher=att_dim,
                        conv_n=3, conv_num_filters=2, num_phonemes=4)
    params = _attention_params(W_state, W_pre, v, K=2, n=3)
    P = O.preprocess(params, attended)
    assert_allclose(P, attended.dot(W_pre))

    s = np.zeros((batch, dim))
    _, w, _, step = O.initial_glimpses(cfg, batch, attended)
    states, glimpses, weights = [], [], []
    for t in range(in_len):
        ctx, w, e, step = O.take_glimpses(cfg, params, attended, P, attended_mask, w, step, s)
        s = O.simple_recurrent_step(s, inputs[t] + ctx.dot(W_dist), W_rec, inputs_mask[t],
                                    activation=lambda z: z)
        states.append(s); glimpses.append(ctx); weights.append(w)
    states, glimpses, weights = map(np.stack, (states, glimpses, weights))
    assert step[0] == in_len
    assert np.all(weights * (1 - attended_mask.T) == 0)
    assert_allclose(weights.sum(), in_len * batch, 1e-5)
    assert_allclose(states.sum(), 113.429, rtol=1e-5)
    assert_allclose(glimpses.sum(), 415.901, rtol=1e-5)
    # recomputing the preprocessed sequence inside take_glimpses (search path) is the same thing
    ctx2, _, _, _ = O.take_glimpses(cfg, params, attended, None, attended_mask, w, step, s)
    ctx1, _, _, _ = O.take_glimpses(cfg, params, attended, P, attended_mask, w, step, s)
    assert_allclose(ctx1, ctx2, rtol=1e-12)


# ---- libs/blocks/tests/bricks/test_sequence_generators.py:197-275 through O.take_glimpses ----

def test_generator_with_attention_freeze_sums_through_take_glimpses():
    inp_dim, inp_len, att_dim, att_len, batch, n_steps = 2, 10, 3, 11, 4, 30
    rng = np.random.RandomState(1234)
    outputs = _rand(rng, (inp_len, batch, inp_dim))
    outputs_mask = _generate_mask(rng, inp_len, batch)
    attended = _rand(rng, (att_len, batch, att_dim))
    attended_mask = _generate_mask(rng, att_len, batch)

    init = np.random.RandomState(1234)
    g = lambda shape: init.normal(0, 0.1, size=shape)
    W_ms = g((inp_dim, inp_dim)); W_mw = g((att_dim, inp_dim))
    W_fork = g((inp_dim, inp_dim)); W_rec = g((inp_dim, inp_dim))
    W_state = g((inp_dim, inp_dim)); W_pre = g((att_dim, inp_dim))
    v = g((inp_dim, 1)); W_dist = g((att_dim, inp_dim))
    ident = lambda z: z

    # E = att_dim = 3 is odd: the config only carries dims for param_shapes, take_glimpses reads shapes from the arrays
    cfg = O.make_config(num_features=3, dims_bidir=[2], dim_dec=inp_dim, dim_matcher=inp_dim,
                        conv_n=2, conv_num_filters=3, num_phonemes=4)
    params = _attention_params(W_state, W_pre, v, K=3, n=2)
    P = O.preprocess(params, attended)

    inputs = outputs.dot(W_fork)
    s = np.zeros((batch, inp_dim))
    w = np.zeros((batch, att_len)); w[:, 0] = 1
    step = np.zeros(batch, dtype=np.int64)
    prev_states, ctxs = [], []
    for i in range(inp_len):
        prev_states.append(s)
        ctx, w, _, step = O.take_glimpses(cfg, params, attended, P, attended_mask, w, step, s)
        s = O.simple_recurrent_step(s, inputs[i] + ctx.dot(W_dist), W_rec, outputs_mask[i], ident)
        ctxs.append(ctx)
    readouts = np.stack(prev_states).dot(W_ms) + np.stack(ctxs).dot(W_mw)
    costs = ((readouts - outputs) ** 2).sum(axis=-1) * outputs_mask
    assert_allclose(costs.sum(), 13.5042, rtol=1e-5)

    s = np.zeros((batch, inp_dim))
    w = np.zeros((batch, att_len)); w[:, 0] = 1
    step = np.zeros(batch, dtype=np.int64)
    S, Y, G, Wts = [], [], [], []
    for i in range(n_steps):
        ctx, w, _, step = O.take_glimpses(cfg, params, attended, P, attended_mask, w, step, s)
        y = s.dot(W_ms) + ctx.dot(W_mw)
        s = O.simple_recurrent_step(s, y.dot(W_fork) + ctx.dot(W_dist), W_rec, None, ident)
        S.append(s); Y.append(y); G.append(ctx); Wts.append(w)
    assert_allclose(np.sum(S), 23.4172, rtol=1e-5)
    assert_allclose(np.sum(Wts), 120.0, rtol=1e-5)
    assert_allclose(np.sum(G), 199.2402, rtol=1e-5)
    assert_allclose(np.sum(Y), -11.6008, rtol=1e-5)


# ---- libs/blocks/tests/bricks/test_sequence_generators.py:96-171 through O.cost_matrix --------

def _integer_generator_model():
    """The reference's SequenceGenerator(Readout(states) + LookupFeedback + SoftmaxEmitter, GRU)
    expressed as a degenerate SpeechRecognizer generator: attention present but disconnected."""
    readout_dim, feedback_dim, dim = 5, 3, 20
    init = np.random.RandomState(1234)
    g = lambda shape: init.normal(0, 0.1, size=shape)
    lookup = g((readout_dim, feedback_dim))
    W_merge = g((dim, readout_dim))
    W_fi = g((feedback_dim, dim))
    W_fg = g((feedback_dim, 2 * dim))
    W_ss = g((dim, dim)); W_su = g((dim, dim)); W_sr = g((dim, dim))

    cfg = O.make_config(num_features=4, dims_bidir=[3], dim_dec=dim, dim_matcher=6, conv_n=2,
                        conv_num_filters=2, num_phonemes=readout_dim, post_merge_dims=[readout_dim],
                        maxout_pieces=1, post_merge_activation="identity",
                        dim_output_embedding=feedback_dim)
    E = O.dim_encoded(cfg)
    junk = np.random.RandomState(7)
    params = O.init_params(cfg, seed=5, weights_std=0.3)       # arbitrary non-zero attention parameters
    params[GEN + "/readout/lookupfeedback/lookuptable.W"] = np.vstack([lookup, junk.normal(size=(1, feedback_dim))])
    params[GEN + "/readout/merge/transform_states.W"] = W_merge
    params[GEN + "/readout/merge/transform_weighted_averages.W"] = np.zeros((E, readout_dim))
    params[GEN + "/readout/post_merge/bias.b"] = np.zeros(readout_dim)
    params[GEN + "/readout/post_merge/mlp/linear_0.W"] = np.eye(readout_dim)
    params[GEN + "/readout/post_merge/mlp/linear_0.b"] = np.zeros(readout_dim)
    params[GEN + "/fork/fork_inputs.W"] = W_fi
    params[GEN + "/fork/fork_inputs.b"] = np.zeros(dim)
    params[GEN + "/fork/fork_gate_inputs.W"] = W_fg
    params[GEN + "/fork/fork_gate_inputs.b"] = np.zeros(2 * dim)
    params[TR + "/transition.state_to_state"] = W_ss
    params[TR + "/transition.state_to_gates"] = np.hstack([W_su, W_sr])
    params[TR + "/transition.initial_state"] = np.zeros(dim)
    params[TR + "/distribute/fork_inputs.W"] = np.zeros((E, dim))
    params[TR + "/distribute/fork_gate_inputs.W"] = np.zeros((E, 2 * dim))
    return cfg, params


def test_integer_generator_freeze_sums_through_cost_matrix():
    cfg, params = _integer_generator_model()
    batch, n_steps, readout_dim = 30, 10, 5
    rng = np.random.RandomState(1234)
    y = rng.randint(readout_dim, size=(n_steps, batch))
    mask = np.ones((n_steps, batch))
    att_rng = np.random.RandomState(3)
    attended = att_rng.normal(size=(7, batch, O.dim_encoded(cfg)))
    attended_mask = np.ones((7, batch))

    r = O.cost_matrix(cfg, params, attended, attended_mask, y, mask, return_all=True)
    costs = r["costs"]
    assert costs.shape == (n_steps, batch)
    assert_allclose(costs.sum(), 482.827, rtol=1e-5)
    assert_allclose(costs.sum(axis=0).mean(), 16.0942, rtol=1e-5)        # generator.cost
    assert_allclose(costs.sum() / mask.sum(), 1.60942, rtol=1e-5)         # per_sequence_element
    assert_allclose(O.batch_cost(costs), 16.0942, rtol=1e-5)              # lvsr/main.py:340-345 == generator.cost here
    # the attention really ran (it is only disconnected): weights are a distribution per step
    assert_allclose(r["weights"].sum(axis=2), 1.0, rtol=1e-9)

    # mask-agnostic cost (:167-171)
    def costs_fun(yy, mm):
        yy = np.asarray(yy); mm = np.asarray(mm, dtype=float)
        return O.cost_matrix(cfg, params, attended[:, :yy.shape[1]], attended_mask[:, :yy.shape[1]], yy, mm)
    c1 = costs_fun([[1], [2]], [[1], [1]])
    c2 = costs_fun([[3, 1], [4, 2], [2, 0]], [[1, 1], [1, 1], [1, 0]])
    assert_allclose(c1.sum(), c2[:, 1].sum(), rtol=1e-5)


def test_compute_states_and_readout_against_closed_forms():
    """O.compute_states == GRU closed form of test_recurrent.py:432-453 plus the Distribute term
    (B/bricks/parallel.py:249-265); O.readout == Merge + Bias + Maxout + Linear literals
    (B/bricks/simple.py:175-181; test_bricks.py Maxout/Linear semantics)."""
    cfg, params = _integer_generator_model()
    rng = np.random.RandomState(0)
    C, E = cfg["dim_dec"], O.dim_encoded(cfg)
    params[TR + "/distribute/fork_inputs.W"] = rng.normal(size=(E, C))
    params[TR + "/distribute/fork_gate_inputs.W"] = rng.normal(size=(E, 2 * C))
    s = rng.normal(size=(3, C)); ctx = rng.normal(size=(3, E))
    a = rng.normal(size=(3, C)); gi = rng.normal(size=(3, 2 * C))
    m = np.array([1.0, 0.0, 1.0])
    got = O.compute_states(cfg, params, s, a, gi, ctx, m)
    Wg, Ws = params[TR + "/transition.state_to_gates"], params[TR + "/transition.state_to_state"]
    gates = O.sigmoid(s.dot(Wg) + gi + ctx.dot(params[TR + "/distribute/fork_gate_inputs.W"]))
    z, r = gates[:, :C], gates[:, C:]
    c = np.tanh((s * r).dot(Ws) + a + ctx.dot(params[TR + "/distribute/fork_inputs.W"]))
    want = c * z + s * (1 - z)
    want = m[:, None] * want + (1 - m[:, None]) * s
    assert_allclose(got, want, rtol=1e-12)
    assert_allclose(got[1], s[1])

    # readout with Maxout(2): adjacent pairs
    cfg2 = O.make_config(num_features=4, dims_bidir=[3], dim_dec=4, conv_n=2, conv_num_filters=2,
                         num_phonemes=3, post_merge_dims=[6], maxout_pieces=2)
    p2 = O.init_params(cfg2, seed=2, weights_std=0.5)
    p2[GEN + "/readout/post_merge/bias.b"] = rng.normal(size=6)
    st = rng.normal(size=(2, 4)); wa = rng.normal(size=(2, 6))
    pre = (st.dot(p2[GEN + "/readout/merge/transform_states.W"]) +
           wa.dot(p2[GEN + "/readout/merge/transform_weighted_averages.W"]) + p2[GEN + "/readout/post_merge/bias.b"])
    mo = np.maximum(pre[:, 0::2], pre[:, 1::2])
    want = mo.dot(p2[GEN + "/readout/post_merge/mlp/linear_0.W"]) + p2[GEN + "/readout/post_merge/mlp/linear_0.b"]
    assert_allclose(O.readout(cfg2, p2, st, wa), want, rtol=1e-12)


# ---- libs/blocks/tests/bricks/test_recurrent.py:455-495 through O.encoder ----------------------

def test_gru_many_steps_masked_through_encoder():
    """The 24-step masked GRU known answer, run by O.encoder itself: the Fork of
    RecurrentWithFork (lvsr/bricks/__init__.py:39-43) is made to reproduce the test's inputs
    (fork_inputs = I, fork_gate_inputs = [-2I | -I] with bias [0.6 | 0.3] gives zi = 2(0.3-x),
    ri = 0.3-x); the backward half doubles as the Bidirectional check of :519-534."""
    rng = np.random.RandomState(1)
    W = rng.normal(0, 1, (3, 3)); Wz = rng.normal(0, 1, (3, 3)); Wr = rng.normal(0, 1, (3, 3))
    x = 0.1 * np.asarray(list(itertools.permutations(range(4))), dtype=float)
    x = np.ones((24, 4, 3)) * x[..., None]
    mask = np.ones((24, 4)); mask[12:24, 3] = 0

    cfg = O.make_config(num_features=3, dims_bidir=[3], dim_dec=4, conv_n=2, conv_num_filters=2, num_phonemes=4)
    params = OrderedDict()
    for d in ("forward", "backward"):
        b = "/recognizer/encoder/bidir0/%s" % d
        params[b + "/gatedrecurrent.state_to_state"] = W
        params[b + "/gatedrecurrent.state_to_gates"] = np.hstack([Wz, Wr])
        params[b + "/gatedrecurrent.initial_state"] = np.zeros(3)
        params[b + "/fork/fork_inputs.W"] = np.eye(3)
        params[b + "/fork/fork_inputs.b"] = np.zeros(3)
        params[b + "/fork/fork_gate_inputs.W"] = np.hstack([-2 * np.eye(3), -np.eye(3)])
        params[b + "/fork/fork_gate_inputs.b"] = np.concatenate([np.full(3, 0.6), np.full(3, 0.3)])
    got, got_mask = O.encoder(cfg, params, x, mask, activation=np.tanh, gate_activation=np.tanh)
    assert got.shape == (24, 4, 6)
    assert_allclose(got_mask, mask)

    ri = 0.3 - x; zi = 2 * ri
    def loop(xs, zis, ris, ms):
        h = np.zeros((25, 4, 3))
        for i in range(1, 25):
            z = np.tanh(h[i - 1].dot(Wz) + zis[i - 1])
            r = np.tanh(h[i - 1].dot(Wr) + ris[i - 1])
            h[i] = np.tanh((r * h[i - 1]).dot(W) + xs[i - 1])
            h[i] = z * h[i] + (1 - z) * h[i - 1]
            h[i] = ms[i - 1, :, None] * h[i] + (1 - ms[i - 1, :, None]) * h[i - 1]
        return h[1:]
    assert_allclose(got[..., :3], loop(x, zi, ri, mask), rtol=1e-6)
    assert_allclose(got[::-1, :, 3:], loop(x[::-1], zi[::-1], ri[::-1], mask[::-1]), rtol=1e-6)

    # Encoder.apply subsampling: x[::k] AFTER the full-rate layer, mask[::k] (lvsr/bricks/__init__.py:75-77)
    cfg2 = dict(cfg, subsample=[3])
    sub, sub_mask = O.encoder(cfg2, params, x, mask, activation=np.tanh, gate_activation=np.tanh)
    assert_allclose(sub, got[::3]); assert_allclose(sub_mask, mask[::3])


def test_one_of_n_feedback_is_a_lookup_of_the_identity():
    """embed_outputs=False (OneOfNFeedback, lvsr/bricks/__init__.py:86-109; exp/wsj/configs/wsj_jan_new.yaml:46) ==
    LookupFeedback whose table is the identity: the pinned cost_matrix path covers it."""
    base = dict(num_features=4, dims_bidir=[3], dim_dec=6, dim_matcher=6, conv_n=2, conv_num_filters=2, num_phonemes=5,
                post_merge_dims=[6], maxout_pieces=2)
    cfg1 = O.make_config(embed_outputs=False, **base)
    cfg2 = O.make_config(dim_output_embedding=6, **base)          # V + 1 = 6
    assert cfg1["dim_feedback"] == 6 and "/recognizer/generator/readout/lookupfeedback/lookuptable.W" not in O.param_shapes(cfg1)
    p1 = O.init_params(cfg1, seed=3, weights_std=0.4)
    p2 = OrderedDict(p1)
    p2["/recognizer/generator/readout/lookupfeedback/lookuptable.W"] = np.eye(6)
    x, m, labels, lm = O.synthetic_batch(cfg1, B=3, T=12, seed=2, label_div=3)
    assert_allclose(O.recognizer_cost(cfg1, p1, x, m, labels, lm), O.recognizer_cost(cfg2, p2, x, m, labels, lm), rtol=1e-12)
