"""Import alias so reference call sites (``import mindaudio``; ``mindaudio.stft``;
``import mindaudio.data.spectrum as spectrum``) resolve to the B200 implementation of the
front-end feature path.  Only that path exists here; see INTEGRATION.md."""
from mindaudio_b200 import *  # noqa: F401,F403
from mindaudio_b200 import data  # noqa: F401
