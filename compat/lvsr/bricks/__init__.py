from .recognizer import SpeechRecognizer  # noqa: F401
