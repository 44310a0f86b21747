import _engine

BurnIn = _engine.pkg.algorithms.BurnIn
