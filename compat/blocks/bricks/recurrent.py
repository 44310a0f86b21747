import _engine

GatedRecurrent = _engine.pkg.GatedRecurrent


class SimpleRecurrent(object):
    """Named so that configs parse; the CUDA path implements GatedRecurrent only (SpeechRecognizer raises)."""


class LSTM(object):
    pass
