"""str-valued stand-ins for the ``mindspore.dataset.audio.utils`` enums the reference coerces its
arguments with (``spectrum.py:592-593,668-671``, ``features.py:331``): both ``"hann"`` and
``WindowType.HANN`` work, without importing mindspore."""
from enum import Enum


class _StrEnum(str, Enum):
    def __str__(self):
        return self.value


class BorderType(_StrEnum):
    CONSTANT = "constant"
    EDGE = "edge"
    REFLECT = "reflect"
    SYMMETRIC = "symmetric"


class MelType(_StrEnum):
    HTK = "htk"
    SLANEY = "slaney"


class NormType(_StrEnum):
    NONE = "none"
    SLANEY = "slaney"


class NormMode(_StrEnum):
    NONE = "none"
    ORTHO = "ortho"


class WindowType(_StrEnum):
    BARTLETT = "bartlett"
    BLACKMAN = "blackman"
    HAMMING = "hamming"
    HANN = "hann"
    KAISER = "kaiser"
