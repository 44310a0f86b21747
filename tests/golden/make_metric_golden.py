"""Generate tests/golden/{metric,config2}_golden.npz: float64 oracle outputs of the EXACT workloads
bench.py times (BASELINE metric config, B=64 x T=1000, L=125) and of BASELINE config 2
(B=32 x T=800, L=100), reduced to what fits in a small fixture:

    costs [L,B] in full; states / weighted averages / weights / energies as random projections
    over their last axis ([L,B] each, projection vectors from RandomState(99)); argmax of the
    weights; the encoder output as a projection [T',B].

Run from the repo root (takes a few minutes):  python tests/golden/make_metric_golden.py
Inputs are NOT stored: bench.synthetic_batch / bench.init_values regenerate them from their seeds.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import lvsr_oracle as O  # noqa: E402

CASES = {"metric": dict(B=64, T=1000, L=125, seed=1234), "config2": dict(B=32, T=800, L=100, seed=1234)}


def projections(E, C, Tp, seed=99):
    rng = np.random.RandomState(seed)
    return dict(pE=rng.normal(size=E), pC=rng.normal(size=C), pT=rng.normal(size=Tp))


def reduce_outputs(r, attended, proj):
    return dict(
        costs=r["costs"].astype(np.float64),
        states_p=r["states"].dot(proj["pC"]),
        wavg_p=r["weighted_averages"].dot(proj["pE"]),
        weights_p=r["weights"].dot(proj["pT"]),
        energies_p=r["energies"].dot(proj["pT"]),
        weights_argmax=r["weights"].argmax(axis=-1).astype(np.int32),
        attended_p=attended.dot(proj["pE"]))


def main():
    for name, c in CASES.items():
        cfg = O.make_config(**bench.NET)
        shapes = O.param_shapes(cfg)
        params = {k: v.astype(np.float64) for k, v in bench.init_values(shapes).items()}
        x, m, labels, lm = bench.synthetic_batch(c["B"], c["T"], 40, c["L"], 32, seed=c["seed"])
        t0 = time.time()
        attended, amask = O.encoder(cfg, params, x.astype(np.float64), m.astype(np.float64))
        r = O.cost_matrix(cfg, params, attended, amask, labels, lm.astype(np.float64), return_all=True)
        proj = projections(attended.shape[2], cfg["dim_dec"], attended.shape[0])
        out = reduce_outputs(r, attended, proj)
        path = os.path.join(ROOT, "tests", "golden", "%s_golden.npz" % name)
        np.savez_compressed(path, **out)
        print("%s: oracle %.1f s -> %s (%.0f KB)" % (name, time.time() - t0, path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
