"""``mindaudio.data.io``: ``read`` (WAV decode feeding the feature path); ``write`` is not on the path."""
from mindaudio_b200.data.io import *  # noqa: F401,F403
