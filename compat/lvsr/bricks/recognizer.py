import _engine

SpeechRecognizer = _engine.pkg.SpeechRecognizer


class SpeechBottom(object):
    """`bottom_class` token of the configs (lvsr/bricks/recognizer.py:105-157); Identity when `dims` is empty."""
