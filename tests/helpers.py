"""Shared helpers for the parity tests: build a CUDA SpeechRecognizer from an oracle config."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402
from oracle import lvsr_oracle as O  # noqa: E402


def package():
    return graft.load_package()


def make_recognizer(cfg, params=None):
    pkg = package()
    act = {"maxout": pkg.Maxout(cfg["maxout_pieces"]), "relu": pkg.Rectifier(), "tanh": pkg.Tanh()}[
        cfg["post_merge_activation"]]
    rec = pkg.SpeechRecognizer(
        input_dims={"recordings": cfg["num_features"]}, input_num_chars={}, eos_label=cfg["eos_label"],
        num_phonemes=cfg["num_phonemes"], dim_dec=cfg["dim_dec"], dims_bidir=cfg["dims_bidir"],
        subsample=cfg["subsample"], conv_n=cfg["conv_n"], conv_num_filters=cfg["conv_num_filters"],
        dim_matcher=cfg["dim_matcher"], post_merge_dims=cfg["post_merge_dims"], post_merge_activation=act,
        dim_output_embedding=cfg["dim_feedback"] if cfg.get("embed_outputs", True) else None,
        embed_outputs=cfg.get("embed_outputs", True), prior=cfg["prior"], energy_normalizer=cfg["energy_normalizer"],
        use_states_for_readout=cfg["use_states_for_readout"],
        max_decoded_length_scale=cfg["max_decoded_length_scale"],
        enc_transition=pkg.GatedRecurrent, dec_transition=pkg.GatedRecurrent, data_prepend_eos=False)
    if params is not None:
        rec.set_parameter_values(params)
    return rec


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max() / max(1e-12, np.abs(want).max()))


SMALL = dict(num_features=40, dims_bidir=[128], subsample=[1], dim_dec=128, conv_n=8, conv_num_filters=10,
             num_phonemes=32, post_merge_dims=[128], maxout_pieces=2)
PYRAMID = dict(num_features=40, dims_bidir=[128, 128, 128], subsample=[1, 2, 2], dim_dec=128, dim_matcher=256,
               conv_n=12, conv_num_filters=10, num_phonemes=32, post_merge_dims=[128], maxout_pieces=2)
WSJ = dict(num_features=40, dims_bidir=[256, 256, 256, 256], subsample=[1, 1, 2, 2], dim_dec=256, dim_matcher=512,
           conv_n=100, conv_num_filters=10, num_phonemes=32, post_merge_dims=[256], maxout_pieces=2)
