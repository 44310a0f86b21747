"""Batched, device-resident beam search (lvsr_search_expand / lvsr_search_advance, BeamSearch.search_many)
against the float64 oracle's line-for-line BeamSearch.search run per utterance: identical token lists."""
import numpy as np
import pytest

from helpers import O, PYRAMID, make_recognizer, package

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _peaky(cfg, seed, gain=10.0, eos_bias=1.0):
    params = O.init_params(cfg, seed=seed, scale=10.0)
    params["/recognizer/generator/readout/post_merge/mlp/linear_0.W"] *= gain
    params["/recognizer/generator/readout/post_merge/mlp/linear_0.b"][cfg["eos_label"]] = eos_bias
    return params


PRIORS = [None, dict(type="window_around_median", before=6, after=8),
          dict(type="expanding", initial_begin=0, initial_end=6, min_speed=0.8, max_speed=2.5)]


@pytest.mark.parametrize("prior", PRIORS, ids=lambda p: "default" if p is None else p["type"])
@pytest.mark.parametrize("beam_size,stop_on,char_discount", [(1, "patience", 0), (5, "patience", 0.0),
                                                             (10, "optimistic_future_cost", 0.1)])
def test_search_many_equals_oracle_per_utterance(prior, beam_size, stop_on, char_discount):
    _torch()
    cfg = O.make_config(prior=prior, max_decoded_length_scale=3.0, **PYRAMID)
    params = _peaky(cfg, 11)
    rng = np.random.RandomState(5)
    utts = [rng.normal(size=(T, cfg["num_features"])) for T in (64, 37, 52, 64, 45, 30)]
    rec = make_recognizer(cfg, params)
    rec.init_beam_search(beam_size)
    got = rec._beam_search.search_many([u.astype(np.float32) for u in utts], cfg["eos_label"],
                                       [int(u.shape[0] / 3.0) for u in utts], stop_on=stop_on,
                                       char_discount=char_discount, raise_on_failure=False)
    n_found = n_hyp = 0
    for u, g in zip(utts, got):
        try:
            want = O.beam_search(cfg, params, u, beam_size, stop_on=stop_on, char_discount=char_discount)
        except O.CandidateNotFoundError:
            assert g is None
            continue
        assert g is not None
        n_found += 1
        n_hyp += len(want[0])
        assert g[0] == want[0]
        assert np.allclose(g[1], want[1], rtol=1e-3, atol=5e-3)
    print("utterances with a result:", n_found, "finished hypotheses compared:", n_hyp)
    if beam_size >= 5:
        assert n_found >= 1 and n_hyp >= 3                     # the case is not degenerate
    # the single-utterance entry point is the same code path with one segment
    one = rec.beam_search({"recordings": utts[0]}, stop_on=stop_on, char_discount=char_discount) if got[0] is not None else None
    if one is not None:
        assert one[0] == got[0][0]
    many = rec.beam_search_many([{"recordings": u} for u in utts[:2]], stop_on=stop_on, char_discount=char_discount,
                                raise_on_failure=False)
    assert [None if r is None else r[0] for r in many] == [None if r is None else r[0] for r in got[:2]]


def test_search_launch_count_is_shared_by_all_utterances():
    """Per step ONE expand + ONE advance whatever the number of utterances (the round-1 loop issued two C calls
    of ~6 launches each per utterance and step, and copied the [width, V] table to the host)."""
    _torch()
    pkg = package()
    cfg = O.make_config(max_decoded_length_scale=4.0, **PYRAMID)
    params = _peaky(cfg, 11)
    rng = np.random.RandomState(6)
    utts = [rng.normal(size=(48, cfg["num_features"])).astype(np.float32) for _ in range(12)]
    rec = make_recognizer(cfg, params)
    rec.init_beam_search(4)
    lib = pkg._lib.load()
    lib.lvsr_launch_count(1)
    rec._beam_search.search_many(utts[:1], cfg["eos_label"], [12], raise_on_failure=False)
    one = lib.lvsr_launch_count(1)
    rec._beam_search.search_many(utts, cfg["eos_label"], [12] * 12, raise_on_failure=False)
    many = lib.lvsr_launch_count(1)
    print("launches: 1 utterance", one, "12 utterances", many)
    assert many <= 1.5 * one                              # not 12x


@pytest.mark.parametrize("stop_on,char_discount", [("patience", 0.0), ("optimistic_future_cost", 0.2)])
def test_native_loop_equals_python_loop(stop_on, char_discount):
    """lvsr_beam_search_many (the loop in C++) returns exactly what the Python mirror of BeamSearch.search returns:
    every finished hypothesis, its per-step costs and the ranking."""
    _torch()
    cfg = O.make_config(prior=dict(type="window_around_mean", before=7, after=7), max_decoded_length_scale=2.5, **PYRAMID)
    params = _peaky(cfg, 4)
    rng = np.random.RandomState(8)
    utts = [rng.normal(size=(T, cfg["num_features"])).astype(np.float32) for T in (60, 33, 48, 25, 57)]
    rec = make_recognizer(cfg, params)
    rec.init_beam_search(6)
    bs = rec._beam_search
    maxl = [int(u.shape[0] / 2.5) for u in utts]
    native = bs.search_many(utts, cfg["eos_label"], maxl, stop_on=stop_on, char_discount=char_discount,
                            raise_on_failure=False, as_arrays=True)
    bs.force_python_loop = True
    python = bs.search_many(utts, cfg["eos_label"], maxl, stop_on=stop_on, char_discount=char_discount,
                            raise_on_failure=False, as_arrays=True)
    bs.force_python_loop = False
    assert len(native) == len(python) == 5
    compared = 0
    for a, b in zip(native, python):
        assert (a is None) == (b is None)
        if a is None:
            continue
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y)
        compared += a[0].shape[1]
    assert compared >= 1
    # a validate_solution_function routes through the Python loop and filters hypotheses
    only_short = bs.search_many(utts[:1], cfg["eos_label"], maxl[:1], stop_on=stop_on, char_discount=char_discount,
                                raise_on_failure=False, validate_solution_function=lambda inputs, seq: len(seq) <= 3)
    if only_short[0] is not None:
        assert all(len(o) <= 2 for o in only_short[0][0])
