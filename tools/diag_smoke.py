"""Stage-by-stage NaN / sentinel diagnosis of the smoke configuration (run plain and under ncu)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
from oracle import lvsr_oracle as O
pkg = g.load_package()
cfg = O.make_config(num_features=40, dims_bidir=[128], subsample=[1], dim_dec=128, conv_n=10,
                    conv_num_filters=10, num_phonemes=32, post_merge_dims=[128], maxout_pieces=2)
params = O.init_params(cfg, seed=1, scale=10.0)
x, m, labels, lm = O.synthetic_batch(cfg, B=4, T=48, seed=7)
ref = O.recognizer_cost(cfg, params, x, m, labels, lm)
rec = pkg.SpeechRecognizer(
    input_dims={"recordings": 40}, input_num_chars={}, eos_label=cfg["eos_label"], num_phonemes=32,
    dim_dec=128, dims_bidir=[128], subsample=[1], conv_n=10, conv_num_filters=10,
    post_merge_dims=[128], post_merge_activation=pkg.Maxout(2),
    enc_transition=pkg.GatedRecurrent, dec_transition=pkg.GatedRecurrent)
rec.set_parameter_values(params)
def nn(t):
    return int(torch.isnan(t).sum().item())
order = os.environ.get("DIAG_ORDER", "cost_first")
if order == "cost_first":
    got = rec.cost(x, m, labels, lm)
    print("cost(): nan", int(np.isnan(got).sum()), "of", got.size, "err", np.nanmax(np.abs(got - ref)) / np.abs(ref).max())
att, attm = rec.encode(x, m)
print("attended nan", nn(att), "mask nan", nn(attm))
P = rec.preprocess(att)
print("P nan", nn(P))
r = rec.cost_matrix(labels, lm, att, attm, return_all=True)
for k, v in r.items():
    print(k, "nan", nn(v), "of", v.numel())
    if nn(v):
        idx = torch.nonzero(torch.isnan(v))
        print("   first nan idx", idx[:4].tolist(), "last", idx[-2:].tolist())
got2 = r["costs"].cpu().numpy()
print("cost_matrix err", np.nanmax(np.abs(got2 - ref)) / np.abs(ref).max())
got = rec.cost(x, m, labels, lm)
print("cost() again: nan", int(np.isnan(got).sum()), "err", np.nanmax(np.abs(got - ref)) / np.abs(ref).max())
