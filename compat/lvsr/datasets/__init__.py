from .npz import Data, NpzAudioDataset  # noqa: F401
