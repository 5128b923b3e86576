from mindaudio_b200.data import *  # noqa: F401,F403
from . import augment, features, io, processing, spectrum  # noqa: F401
