"""attention-lvcsr_b200 -- B200-native hot path of rizar/attention-lvcsr.

The directory name contains a hyphen (it is the contract's name); import it through
``__graft_entry__.load_package()`` which registers it as ``attention_lvcsr_b200``.

Only what the path needs lives here: ``csrc/`` (sm_100a kernels + the C ABI of
include/lvsr_b200.h) and the host-side mirror of the reference's operator surface
(``SpeechRecognizer``, ``BeamSearch``, initialisation/config tokens).
"""
from . import _lib  # noqa: F401
from .bricks import (Constant, GatedRecurrent, Identity, IsotropicGaussian, Maxout,  # noqa: F401
                     Orthogonal, Rectifier, Tanh, Uniform)
from . import algorithms  # noqa: F401
from .algorithms import (AdaDelta, BurnIn, CompositeRule, GradientDescent, Momentum, RemoveNotFinite,  # noqa: F401
                         Restrict, Scale, StepClipping, VariableClipping, step_rule_from_config)
from .recognizer import SpeechRecognizer  # noqa: F401
from .search import BeamSearch, CandidateNotFoundError  # noqa: F401

__all__ = ["GradientDescent", "CompositeRule", "StepClipping", "Momentum", "AdaDelta", "VariableClipping", "Restrict",
           "RemoveNotFinite", "BurnIn", "Scale", "step_rule_from_config", "SpeechRecognizer", "BeamSearch", "CandidateNotFoundError", "Maxout", "Rectifier", "Tanh",
           "Identity", "GatedRecurrent", "IsotropicGaussian", "Constant", "Orthogonal", "Uniform"]
