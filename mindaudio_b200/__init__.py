"""mindaudio_b200 -- MindAudio's front-end feature path (spectrum / features / CMVN) on B200.

Flat re-exports mirror ``mindaudio/__init__.py`` -> ``mindaudio/data/__init__.py`` of the
reference, so ``mindaudio_b200.stft`` / ``.magphase`` / ``.fbank`` work like ``mindaudio.stft``.
Importing this package does not touch CUDA (fork safe); the first op creates the context.
"""
from ._enums import BorderType, MelType, NormMode, NormType, WindowType  # noqa: F401
from ._lib import MafeError  # noqa: F401
from .data import augment, cmvn, collate, features, io, masking, processing, spectrum  # noqa: F401
from .data.augment import pitch_shift, speed_perturb, time_stretch  # noqa: F401
from .data.collate import *  # noqa: F401,F403
from .data.io import *  # noqa: F401,F403
from .data.masking import *  # noqa: F401,F403
from .data.processing import *  # noqa: F401,F403
from .data.cmvn import *  # noqa: F401,F403
from .data.features import *  # noqa: F401,F403
from .data.spectrum import *  # noqa: F401,F403
from .data.features import mel_and_fbank, mel_and_mfcc  # noqa: F401  (extensions: two outputs of one transform)
from .frontend import FbankPipeline, compute_fbank_feats, ds2_features  # noqa: F401

__version__ = "0.1.0"
