"""``mindaudio/utils/mask.py`` pieces on the feature path."""
from mindaudio_b200.data.collate import make_pad_mask  # noqa: F401
