"""``mindaudio/utils/load_files.py``: ``load_cmvn`` (global CMVN mean / istd from the JSON statistics)."""
from mindaudio_b200.data.cmvn import load_cmvn  # noqa: F401
