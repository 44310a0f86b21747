"""Model-object boundary: initialize() ordering, Blocks-format checkpoints, pickling,
sample / generate, named inputs."""
import io
import pickle
import tarfile

import numpy as np
import pytest

from helpers import O, PYRAMID, SMALL, make_recognizer, package, rel_err

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def test_initialize_walks_the_brick_tree_like_blocks():
    """One RandomState shared down the tree, children first, Linear draws bias then W, recurrent bricks take
    rec_weights_init for all three matrices (B/bricks/interfaces.py:157-200, recognizer.py:363-373): equals
    oracle.init_params, whose order the reference's frozen sums pin."""
    _torch()
    pkg = package()
    cfg = O.make_config(**PYRAMID)
    rec = make_recognizer(cfg)
    rec.weights_init = pkg.IsotropicGaussian(0.01)
    rec.biases_init = pkg.Constant(0.0)
    rec.rec_weights_init = pkg.Orthogonal()
    rec.initial_states_init = pkg.IsotropicGaussian(0.001)
    rec.initialize(seed=1)
    want = O.init_params(cfg, seed=1)
    got = rec.get_parameter_values()
    assert list(got) == list(want)
    for k in want:
        assert np.allclose(got[k], want[k].astype(np.float32), rtol=1e-6, atol=1e-9), k


def test_blocks_checkpoint_round_trip_and_lenient_loading(tmp_path, caplog):
    _torch()
    cfg = O.make_config(**SMALL)
    params = O.init_params(cfg, seed=3, scale=10.0)
    rec = make_recognizer(cfg, params)
    path = str(tmp_path / "model.tar")
    rec.save_params(path)
    # the archive is what blocks.serialization.load_parameters reads: member `_parameters`, npz keys with '|' for '/'
    with tarfile.open(path) as tar:
        assert tar.getnames() == ["_parameters"]
        npz = np.load(io.BytesIO(tar.extractfile("_parameters").read()))
        assert all("/" not in k and k.startswith("|recognizer|") for k in npz.files)
        assert set(k.replace("|", "/") for k in npz.files) == set(params)
    rec2 = make_recognizer(cfg)
    info = rec2.load_params(path)
    assert info == dict(unknown=[], missing=[])
    x, m, labels, lm = O.synthetic_batch(cfg, B=3, T=20, seed=1)
    assert np.array_equal(rec.cost(x, m, labels, lm), rec2.cost(x, m, labels, lm))
    # unknown names are logged, missing parameters keep their values (Model.set_parameter_values semantics)
    vals = {k.replace("/", "|"): v for k, v in rec.get_parameter_values().items()}
    dropped = "|recognizer|generator|readout|post_merge|bias.b"
    del vals[dropped]
    vals["|recognizer|not|a|brick.W"] = np.zeros((2, 2), dtype=np.float32)
    p2 = str(tmp_path / "partial.npz")
    np.savez(p2, **vals)
    rec3 = make_recognizer(cfg)
    with caplog.at_level("ERROR"):
        info = rec3.load_params(p2)
    assert info["unknown"] == ["/recognizer/not/a/brick.W"] and info["missing"] == [dropped.replace("|", "/")]
    assert "unknown parameter names" in caplog.text and "missing values for parameters" in caplog.text


def test_pickle_round_trip_keeps_the_model():
    _torch()
    cfg = O.make_config(**SMALL)
    rec = make_recognizer(cfg, O.init_params(cfg, seed=3, scale=10.0))
    rec.init_beam_search(3)
    x, m, labels, lm = O.synthetic_batch(cfg, B=2, T=24, seed=2)
    want = rec.cost(x, m, labels, lm)
    blob = pickle.dumps(rec)
    rec2 = pickle.loads(blob)
    assert rec2._handle is not None and rec2._beam_search is None       # handle rebuilt, compiled search dropped
    assert np.array_equal(rec2.cost(x, m, labels, lm), want)
    assert list(rec2.inputs.keys()) == ["recordings"] and rec2.labels.name == "labels"


def test_generate_argmax_and_sampling_distribution():
    torch = _torch()
    cfg = O.make_config(**SMALL)
    params = O.init_params(cfg, seed=3, scale=10.0)
    params["/recognizer/generator/readout/post_merge/mlp/linear_0.W"] *= 8.0
    rec = make_recognizer(cfg, params)
    x, m, _, _ = O.synthetic_batch(cfg, B=3, T=40, seed=7)
    # arg-max emission == the oracle's greedy generate
    att, attm = O.encoder(cfg, params, x, m)
    ys, costs, _ = O.generate_greedy(cfg, params, att, attm, 6)
    g = rec.generate(x, m, n_steps=6, sample=False)
    assert np.array_equal(g["outputs"], ys)
    assert rel_err(g["costs"], costs) < 1e-3
    # sampling: 512 copies of one utterance, the first symbol follows softmax(readout) of the oracle
    one = np.repeat(x[:, :1], 512, axis=1)
    s = rec.generate(one, None, n_steps=2, sample=True, seed=5)
    att1, attm1 = O.encoder(cfg, params, x[:, :1], None)
    st = O.initial_states(cfg, params, 1, att1)
    p = np.exp(-O.logprobs_computer(cfg, params, att1, attm1, st))[0]
    freq = np.bincount(s["outputs"][0], minlength=cfg["num_phonemes"]) / 512.0
    assert 0.5 * np.abs(freq - p).sum() < 0.12                     # total variation, 512 draws over 32 symbols
    assert np.allclose(s["costs"][0], -np.log(p[s["outputs"][0]]), rtol=1e-3, atol=1e-3)
    s2 = rec.generate(one, None, n_steps=2, sample=True, seed=5)
    assert np.array_equal(s["outputs"], s2["outputs"])             # seeded: reproducible
    out = rec.sample({"recordings": x[:, 0]}, n_steps=5)
    assert out.shape == (5, 1) and out.dtype == np.int64
    assert rec.get_cost_graph(batch=False)(x[:, 0], ys[:, 0]).shape == (6,)


def test_compat_entry_points_train_search_sample(tmp_path, capsys):
    """compat/lvsr.main.{train_multistage, search, sample} -- what the reference's bin/run.py dispatches to --
    on a toy experiment: multi-stage training writes Blocks-format checkpoints, search prints the reference's
    report lines and decodes with the trained parameters."""
    _torch()
    import sys
    from compat_helpers import COMPAT, write_experiment
    if COMPAT not in sys.path:
        sys.path.insert(0, COMPAT)
    import lvsr.config as C
    import lvsr.main as M
    exp = write_experiment(tmp_path)
    cfg = C.Configuration(exp["child"], None, [])
    cfg["cmd_args"] = {}
    save = str(tmp_path / "run")
    M.train_multistage(cfg, save, "", None, "")
    import os
    import tarfile
    assert sorted(os.listdir(save)) == ["main.tar", "pretraining.tar"]
    with tarfile.open(os.path.join(save, "main.tar")) as tar:
        assert tar.getnames() == ["_parameters"]
    capsys.readouterr()
    single = C.Configuration(exp["base"], None, [("monitoring.search.beam_size", "2")])
    decoded = str(tmp_path / "decoded.txt")
    M.search(single, None, os.path.join(save, "main.tar"), "valid", None, None, decoded, False, 1)
    out = capsys.readouterr().out
    for line in ("Utterance 0 (valid_000)", "Groundtruth:", "Groundtruth cost:", "Decoding took:", "Beam search cost:",
                 "Recognized:", "CER:", "Average CER:"):
        assert line in out, (line, out[-1500:])
    assert len(open(decoded).read().strip().splitlines()) == 3
    M.sample(single, None, os.path.join(save, "main.tar"), "valid")
    assert "Utterance 2" in capsys.readouterr().out
    with pytest.raises(NotImplementedError):
        M.init_norm(single, "x")
