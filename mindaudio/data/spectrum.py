from mindaudio_b200.data.spectrum import *  # noqa: F401,F403
from mindaudio_b200.data.spectrum import _pad_center, _pad_shape, frame  # noqa: F401  (processing.py:9, augment.py:11 import these)
