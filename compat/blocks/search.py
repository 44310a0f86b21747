import _engine

BeamSearch = _engine.pkg.BeamSearch
CandidateNotFoundError = _engine.pkg.CandidateNotFoundError
