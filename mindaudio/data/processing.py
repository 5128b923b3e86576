"""``mindaudio.data.processing`` pieces on the feature path."""
from mindaudio_b200.data.processing import *  # noqa: F401,F403
