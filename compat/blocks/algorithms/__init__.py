import _engine

_a = _engine.pkg.algorithms
GradientDescent = _a.GradientDescent
StepRule, CompositeRule, Scale, BasicMomentum, Momentum = _a.StepRule, _a.CompositeRule, _a.Scale, _a.BasicMomentum, _a.Momentum
AdaDelta, StepClipping, VariableClipping, Restrict, RemoveNotFinite = (_a.AdaDelta, _a.StepClipping, _a.VariableClipping,
                                                                      _a.Restrict, _a.RemoveNotFinite)
