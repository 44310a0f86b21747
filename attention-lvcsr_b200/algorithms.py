"""Training algorithm surface of the reference over the CUDA training step.

Mirrors ``blocks.algorithms`` as lvsr uses it (lvsr/main.py:480-519): the step rules are
configuration objects with the reference's names and constructor arguments,
``GradientDescent(...).process_batch(batch)`` runs one update
(libs/blocks/blocks/algorithms/__init__.py:244-256,284-287).  No symbolic graph exists: the chain
is mapped onto ``lvsr_train_config`` and executed by ``lvsr_train_cost_and_grads`` /
``lvsr_train_apply_updates`` (include/lvsr_b200.h).  A chain the CUDA step does not implement raises
``NotImplementedError`` instead of being approximated.

Data parallelism (the reference is single-device): with ``torch.distributed``
initialised, every rank computes the gradient SUM of its utterance shard, ONE all-reduce carries the
flat gradient buffer together with the local batch size and cost, and every replica applies the same
update with 1 / (global batch size).
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib


class StepRule(object):
    pass


class StepClipping(StepRule):
    """B/algorithms/__init__.py:610-643."""

    def __init__(self, threshold=None):
        self.threshold = threshold


class Scale(StepRule):
    def __init__(self, learning_rate=1.0):
        self.learning_rate = learning_rate


class BasicMomentum(StepRule):
    def __init__(self, momentum=0.0):
        self.momentum = momentum


class Momentum(StepRule):
    """Scale(learning_rate) then BasicMomentum(momentum), B/algorithms/__init__.py:431-461."""

    def __init__(self, learning_rate=1.0, momentum=0.0):
        self.learning_rate = learning_rate
        self.momentum = momentum


class AdaDelta(StepRule):
    """:464-516."""

    def __init__(self, decay_rate=0.95, epsilon=1e-6):
        if not 0.0 <= decay_rate <= 1.0:
            raise ValueError("decay rate needs to be in [0, 1]")
        self.decay_rate = decay_rate
        self.epsilon = epsilon


class VariableClipping(StepRule):
    """:646-720; the CUDA step implements axis=0 (the only use in lvsr/main.py:503-505)."""

    def __init__(self, threshold, axis=None):
        self.threshold = threshold
        self.axis = axis


class Restrict(StepRule):
    """:864-893.  ``variables``: parameter names, or the string "WEIGHT" for every parameter with the
    WEIGHT role (what lvsr/main.py:492 selects)."""

    def __init__(self, step_rule, variables="WEIGHT"):
        self.step_rule = step_rule
        self.variables = variables


class RemoveNotFinite(StepRule):
    """:829-861.  lvsr passes scaler=0.0 (lvsr/main.py:516): a parameter whose step is not finite is ZEROED."""

    def __init__(self, scaler=1):
        self.scaler = scaler


class BurnIn(StepRule):
    """lvsr/algorithms.py:19-43."""

    def __init__(self, num_steps=0):
        self.num_steps = num_steps


class CompositeRule(StepRule):
    def __init__(self, components):
        self.components = list(components)


def step_rule_from_config(train_conf, reg_conf=None):
    """The CompositeRule lvsr/main.py:480-516 builds from config['training'] / config['regularization']."""
    reg_conf = reg_conf or {}
    rules = [StepClipping(train_conf["gradient_threshold"])]
    names = train_conf.get("rules", ["momentum"])
    if "momentum" in names:
        rules.append(Momentum(train_conf["scale"], train_conf["momentum"]))
    if "adadelta" in names:
        rules.append(AdaDelta(train_conf["decay_rate"], train_conf["epsilon"]))
    if reg_conf.get("max_norm", False) > 0:
        rules.append(Restrict(VariableClipping(reg_conf["max_norm"], axis=0), "WEIGHT"))
    rules.append(RemoveNotFinite(0.0))
    if train_conf.get("burn_in_steps", 0):
        rules.append(BurnIn(num_steps=train_conf["burn_in_steps"]))
    return CompositeRule(rules)


LvsrTrainConfig = _lib.LvsrTrainConfig


def _to_train_config(step_rule, decay=0.0):
    """Accepts exactly the chain shapes lvsr/main.py can build, in that order."""
    comps = step_rule.components if isinstance(step_rule, CompositeRule) else [step_rule]
    cfg = LvsrTrainConfig()
    cfg.gradient_threshold = 0.0
    cfg.decay = float(decay)
    stage = 0          # 0 clipping, 1 momentum, 2 adadelta, 3 max-norm, 4 remove-not-finite, 5 burn-in
    seen_rnf = False
    for r in comps:
        if isinstance(r, StepClipping) and stage <= 0:
            cfg.gradient_threshold = float(r.threshold or 0.0)
            stage = 1
        elif isinstance(r, Momentum) and stage <= 1:
            cfg.use_momentum, cfg.scale, cfg.momentum = 1, float(r.learning_rate), float(r.momentum)
            stage = 2
        elif isinstance(r, Scale) and stage <= 1:
            cfg.use_momentum, cfg.scale, cfg.momentum = 1, float(r.learning_rate), 0.0
            stage = 2
        elif isinstance(r, AdaDelta) and stage <= 2:
            cfg.use_adadelta, cfg.decay_rate, cfg.epsilon = 1, float(r.decay_rate), float(r.epsilon)
            stage = 3
        elif isinstance(r, Restrict) and stage <= 3 and isinstance(r.step_rule, VariableClipping) and \
                r.variables == "WEIGHT" and r.step_rule.axis == 0:
            cfg.max_norm = float(r.step_rule.threshold)
            stage = 4
        elif isinstance(r, RemoveNotFinite) and stage <= 4:
            if r.scaler != 0.0:
                raise NotImplementedError("RemoveNotFinite(scaler=%r): the CUDA step implements scaler=0.0 "
                                          "(lvsr/main.py:516)" % (r.scaler,))
            seen_rnf = True
            stage = 5
        elif isinstance(r, BurnIn) and stage <= 5:
            cfg.burn_in_steps = int(r.num_steps)
            stage = 6
        else:
            raise NotImplementedError("step rule chain %s is not one lvsr/main.py:480-516 builds"
                                      % [type(c).__name__ for c in comps])
    if not seen_rnf:
        raise NotImplementedError("the CUDA step always applies RemoveNotFinite(0.0) (lvsr/main.py:516): add it to the chain")
    return cfg


def allreduce_step_buffer(buf, n, local_batch, local_cost, dist):
    """The ONE collective of a data-parallel training step: sum over ranks of
    [flat gradient (n floats) | local batch size | local cost sum].  ``buf`` is a 1-D float32 tensor of at
    least n + 2 elements on any device ``dist`` can reduce (NCCL: the GPU buffer the backward pass wrote;
    gloo: a CPU tensor in the host-side tests).  Returns (global batch size, global cost sum) as 0-d tensors
    that live in ``buf`` -- reading them on the host synchronises."""
    buf[n] = float(local_batch)
    buf[n + 1:n + 2] = local_cost.reshape(1).to(buf.dtype) if hasattr(local_cost, "reshape") else float(local_cost)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf[n], buf[n + 1]


class GradientDescent(object):
    """``GradientDescent(cost=..., parameters=..., step_rule=...)`` of the reference with the recognizer in
    place of the symbolic cost (there is no graph to differentiate: the backward pass is part of the library).

    recognizer: attention_lvcsr_b200.SpeechRecognizer;  step_rule: CompositeRule as built by lvsr/main.py;
    decay: config['regularization']['decay'] (lvsr/main.py:419-421)."""

    def __init__(self, recognizer=None, step_rule=None, decay=0.0, cost=None, parameters=None, gradients=None,
                 on_unused_sources="warn", **kwargs):
        if recognizer is None:
            raise ValueError("GradientDescent needs the recognizer (no symbolic cost exists in the CUDA path)")
        self.recognizer = recognizer
        self.step_rule = step_rule if step_rule is not None else CompositeRule([Scale(), RemoveNotFinite(0.0)])
        self._tc = _to_train_config(self.step_rule, decay)
        self.on_unused_sources = on_unused_sources
        self._grads = None
        self._buf = None
        self._cost = None
        self.equal_shards = True
        self.last_cost = None
        self.last_batch_size = None

    SOURCES = ("recordings", "recordings_mask", "labels", "labels_mask")

    def initialize(self):
        rec = self.recognizer
        torch = rec._torch()
        lib, h = _lib.load(), rec._require_ready()
        n = int(lib.lvsr_model_flat_size(h))
        # [flat gradient | local batch size | local cost sum | padding]: one buffer, one all-reduce
        self._buf = torch.zeros((n + 64,), dtype=torch.float32, device=rec.device)
        self._cost = torch.zeros((1,), dtype=torch.float32, device=rec.device)
        self._n = n
        _lib.check(lib.lvsr_train_reset(h))

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist, dist.get_world_size()
        return None, 1

    def cost_and_gradients(self, batch):
        """(cost, {parameter name: gradient}) of sum(cost_matrix)/B for this batch on this GPU (no update)."""
        rec = self.recognizer
        if self._buf is None:
            self.initialize()
        B = self._forward_backward(batch, None)
        import ctypes as C_
        lib, h = _lib.load(), rec._require_ready()
        flat = self._buf[:self._n].cpu().numpy()
        out = OrderedDict()
        for i, (name, shape) in enumerate(rec.parameter_shapes().items()):
            off, cnt = C_.c_int64(), C_.c_int64()
            _lib.check(lib.lvsr_model_param_offset(h, i, C_.byref(off), C_.byref(cnt)))
            out[name] = flat[off.value:off.value + cnt.value].reshape(shape).copy()
        return float(self._cost.item()), out

    def _forward_backward(self, batch, gscale):
        rec = self.recognizer
        torch = rec._torch()
        lib, h = _lib.load(), rec._require_ready()
        batch = dict(batch)
        unknown = set(batch) - set(self.SOURCES)
        if unknown and self.on_unused_sources == "raise":
            raise ValueError("mismatch of variable names and data sources: %s" % sorted(unknown))
        missing = [s for s in ("recordings", "labels") if s not in batch]
        if missing:
            raise ValueError("Didn't find all sources: %s" % missing)
        x = rec._dev(batch["recordings"], torch.float32)
        m = rec._dev(batch.get("recordings_mask"), torch.float32)
        rec._check_labels(batch["labels"])
        y = rec._dev(batch["labels"], torch.int64)
        ym = rec._dev(batch.get("labels_mask"), torch.float32)
        if x.dim() != 3 or y.dim() != 2:
            raise ValueError("recordings [T,B,F] and labels [L,B] expected")
        T, B, F = x.shape
        L = y.shape[0]
        if F != rec.net["num_features"] or y.shape[1] != B or (m is not None and tuple(m.shape) != (T, B)) or \
                (ym is not None and tuple(ym.shape) != (L, B)):
            raise ValueError("batch shapes disagree: recordings %s mask %s labels %s labels_mask %s" % (
                tuple(x.shape), None if m is None else tuple(m.shape), tuple(y.shape), None if ym is None else tuple(ym.shape)))
        gs = (1.0 / B) if gscale is None else gscale
        _lib.check(lib.lvsr_train_cost_and_grads(
            h, x.data_ptr(), None if m is None else m.data_ptr(), y.data_ptr(), None if ym is None else ym.data_ptr(),
            T, B, L, float(gs), self._cost.data_ptr(), self._buf.data_ptr(), rec._stream()))
        return B

    def process_batch(self, batch):
        """One update (B/algorithms/__init__.py:284-287): parameters change in place on the device."""
        rec = self.recognizer
        if self._buf is None:
            self.initialize()
        lib, h = _lib.load(), rec._require_ready()
        dist, world = self._world()
        if world == 1:
            B = self._forward_backward(batch, None)          # grads already carry 1/B
            _lib.check(lib.lvsr_train_apply_updates(h, self._buf.data_ptr(), 1.0, C.byref(self._tc), rec._stream()))
            self.last_batch_size = B
            self.last_cost = self._cost                       # device scalar; .item() synchronises
            return
        B = self._forward_backward(batch, 1.0)                # gradient SUM over the local utterances
        bg_dev, cost_dev = allreduce_step_buffer(self._buf, self._n, B, self._cost, dist)   # the ONE collective of the step
        # the global batch size has to reach the host to become a kernel argument; every rank knows its own B and
        # shards are equal-sized in the data-parallel loop, so the common case needs no synchronisation
        Bg = B * world if self.equal_shards else int(round(float(bg_dev.item())))
        _lib.check(lib.lvsr_train_apply_updates(h, self._buf.data_ptr(), 1.0 / Bg, C.byref(self._tc), rec._stream()))
        self.last_batch_size = Bg
        self.last_cost = cost_dev / Bg

    def total_gradient_norm(self):
        lib, h = _lib.load(), self.recognizer._require_ready()
        v = C.c_float()
        _lib.check(lib.lvsr_train_gradient_norm(h, C.byref(v)))
        return float(v.value)
