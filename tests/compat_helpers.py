"""A tiny experiment (YAML config in the reference's style + flat .npz data) for the compat/ entry points."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "compat")

BASE_YAML = """
data:
    path: {npz}
    add_eos: True
    prepend_eos: False
    batch_size: 4
    sort_k_batches: 2
initialization:
    /recognizer:
        weights_init:
            !!python/object/apply:blocks.initialization.IsotropicGaussian [0.1]
        biases_init:
            !!python/object/apply:blocks.initialization.Constant [0.0]
        rec_weights_init:
            !!python/object/apply:blocks.initialization.Orthogonal []
        initial_states_init:
            !!python/object/apply:blocks.initialization.IsotropicGaussian [0.001]
net:
    bottom:
        activation: !!python/object/apply:blocks.bricks.Rectifier []
        dims: []
    enc_transition: !!python/name:blocks.bricks.recurrent.GatedRecurrent
    dims_bidir: [128]
    subsample: [1]
    dims_top: []
    dec_transition: !!python/name:blocks.bricks.recurrent.GatedRecurrent
    dec_stack: 1
    dim_dec: 128
    attention_type: content_and_conv
    conv_n: 8
    conv_num_filters: 4
    post_merge_dims: [128]
    post_merge_activation: !!python/object/apply:blocks.bricks.Maxout [2]
    use_states_for_readout: True
    max_decoded_length_scale: 2.0
    criterion:
        name: log_likelihood
    lm: {{}}
regularization:
    max_norm: 1
training:
    rules: [momentum, adadelta]
    scale: 1.0
    momentum: 0.0
    decay_rate: 0.95
    epsilon: 1e-6
    gradient_threshold: 10.0
    num_epochs: 1
monitoring:
    search:
        beam_size: 3
        char_discount: 0.0
        round_to_inf: 1000000000.0
        stop_on: patience
"""

CHILD_YAML = """
parent: {base}
net:
    dim_dec: 128
training:
    num_batches: 3
stages:
    pretraining:
        number: 0
        training:
            num_batches: 2
    main:
        number: 1
        training:
            scale: 0.5
"""


def write_experiment(tmp_path, n_train=10, n_valid=3, F=40, V=12):
    rng = np.random.RandomState(0)
    arrays = dict(num_labels=np.int64(V), characters=np.array(list("abcdefghijk") + ["$"]))
    for part, n in (("train", n_train), ("valid", n_valid)):
        lens = rng.randint(24, 49, size=n)
        llens = np.maximum(2, lens // 8)
        arrays[part + "_features"] = rng.normal(size=(int(lens.sum()), F)).astype(np.float32)
        arrays[part + "_feature_offsets"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        arrays[part + "_labels"] = rng.randint(0, V - 1, size=int(llens.sum())).astype(np.int64)
        arrays[part + "_label_offsets"] = np.concatenate([[0], np.cumsum(llens)]).astype(np.int64)
        arrays[part + "_uttids"] = np.array(["%s_%03d" % (part, i) for i in range(n)])
    npz = os.path.join(str(tmp_path), "toy.npz")
    np.savez(npz, **arrays)
    base = os.path.join(str(tmp_path), "base.yaml")
    with open(base, "w") as f:
        f.write(BASE_YAML.format(npz=npz))
    child = os.path.join(str(tmp_path), "child.yaml")
    with open(child, "w") as f:
        f.write(CHILD_YAML.format(base=base))
    return dict(npz=npz, base=base, child=child)
