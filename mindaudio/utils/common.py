"""``mindaudio/utils/common.py`` pieces on the feature path: ``pad_sequence`` (GPU for float feature matrices)."""
from mindaudio_b200.data.collate import IGNORE_ID, pad_sequence  # noqa: F401
