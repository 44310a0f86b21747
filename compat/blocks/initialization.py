import _engine

IsotropicGaussian = _engine.pkg.IsotropicGaussian
Constant = _engine.pkg.Constant
Orthogonal = _engine.pkg.Orthogonal
Uniform = _engine.pkg.Uniform
