"""Stub ``mindspore.dataset.audio.utils``: str-valued enums + ``create_dct``."""
from enum import Enum

import numpy as np


class BorderType(str, Enum):
    CONSTANT = "constant"
    EDGE = "edge"
    REFLECT = "reflect"
    SYMMETRIC = "symmetric"


class MelType(str, Enum):
    HTK = "htk"
    SLANEY = "slaney"


class NormType(str, Enum):
    NONE = "none"
    SLANEY = "slaney"


class NormMode(str, Enum):
    NONE = "none"
    ORTHO = "ortho"


class WindowType(str, Enum):
    BARTLETT = "bartlett"
    BLACKMAN = "blackman"
    HAMMING = "hamming"
    HANN = "hann"
    KAISER = "kaiser"


class ScaleType(str, Enum):
    MAGNITUDE = "magnitude"
    POWER = "power"


def create_dct(n_mfcc, n_mels, norm=NormMode.NONE):
    from oracle import restated
    return restated.create_dct(n_mfcc, n_mels, NormMode(norm).value).astype(np.float32)
