import _engine
from . import recurrent  # noqa: F401

Maxout = _engine.pkg.Maxout
Rectifier = _engine.pkg.Rectifier
Tanh = _engine.pkg.Tanh
Identity = _engine.pkg.Identity
