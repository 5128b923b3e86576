from mindaudio_b200.data.features import *  # noqa: F401,F403
