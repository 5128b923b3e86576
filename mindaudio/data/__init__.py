from mindaudio_b200.data import *  # noqa: F401,F403
from . import features, spectrum  # noqa: F401
