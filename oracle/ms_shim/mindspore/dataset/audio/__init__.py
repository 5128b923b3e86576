"""Stub ``mindspore.dataset.audio``: eager callables delegating to ``oracle.restated``."""
import numpy as np

from . import utils  # noqa: F401
from .utils import BorderType, MelType, NormType, WindowType


def _v(e):
    return getattr(e, "value", e)


class Spectrogram:
    def __init__(self, n_fft=400, win_length=None, hop_length=None, pad=0, window=WindowType.HANN,
                 power=2.0, normalized=False, center=True, pad_mode=BorderType.REFLECT, onesided=True):
        self.kw = dict(n_fft=n_fft, win_length=win_length, hop_length=hop_length, pad=pad,
                       window=_v(window), power=power, normalized=normalized, center=center,
                       pad_mode=_v(pad_mode), onesided=onesided)

    def __call__(self, x):
        from oracle import restated
        return restated.spectrogram(np.asarray(x), **self.kw)


class MelScale:
    def __init__(self, n_mels=128, sample_rate=16000, f_min=0.0, f_max=None, n_stft=201,
                 norm=NormType.NONE, mel_type=MelType.HTK):
        self.kw = dict(n_mels=n_mels, sample_rate=sample_rate, f_min=f_min, f_max=f_max,
                       n_stft=n_stft, norm=_v(norm), mel_type=_v(mel_type))

    def __call__(self, x):
        from oracle import restated
        return restated.melscale(np.asarray(x), **self.kw)


class ComputeDeltas:
    def __init__(self, win_length=5, pad_mode=BorderType.EDGE):
        self.win_length, self.pad_mode = win_length, _v(pad_mode)

    def __call__(self, x):
        from oracle import restated
        return restated.compute_deltas(np.asarray(x), self.win_length, self.pad_mode)


class Magphase:
    def __init__(self, power=1.0):
        self.power = power

    def __call__(self, x):
        from oracle import restated
        return restated.magphase_real(np.asarray(x), self.power)


class ComplexNorm:
    def __init__(self, power=1.0):
        self.power = power

    def __call__(self, x):
        x = np.asarray(x)
        return (x[..., 0] ** 2 + x[..., 1] ** 2) ** (0.5 * self.power)


class Angle:
    def __call__(self, x):
        x = np.asarray(x)
        return np.arctan2(x[..., 1], x[..., 0])


class SpectralCentroid:
    def __init__(self, sample_rate, n_fft=400, win_length=None, hop_length=None, pad=0,
                 window=WindowType.HANN):
        self.sr = sample_rate
        self.kw = dict(n_fft=n_fft, win_length=win_length, hop_length=hop_length, pad=pad,
                       window=_v(window), power=1.0)

    def __call__(self, x):
        from oracle import restated
        spec = restated.spectrogram(np.asarray(x), **self.kw)
        freqs = np.linspace(0, self.sr // 2, spec.shape[-2]).reshape((-1, 1))
        return (freqs * spec).sum(axis=-2) / spec.sum(axis=-2)
