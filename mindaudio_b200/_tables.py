"""Host-side constant tables handed to the C ABI at plan creation: analysis windows, mel
filterbanks, DCT-II matrices.  Computed once per configuration in float64 and rounded to
float32 for the device.  (Constants, not the hot path: the per-sample arithmetic is all CUDA.)"""
from __future__ import annotations

import functools
import math

import numpy as np
from scipy.signal import get_window as _get_window

from ._enums import MelType, NormMode, NormType, WindowType


def analysis_window(window, win_length, n_fft, coerce=False):
    """Periodic window of ``win_length`` zero-padded symmetrically to ``n_fft``
    (``spectrum.py:173-175`` / the Spectrogram op).  ``coerce``: validate against WindowType the
    way the MindSpore-backed wrappers do (``spectrum.py:592``)."""
    if coerce:
        window = WindowType(window).value
        spec = ("kaiser", 12.0) if window == "kaiser" else window
    else:
        spec = getattr(window, "value", window)
    w = _get_window(spec, win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    if lpad < 0:
        raise ValueError("Target size ({:d}) must be at least input size ({:d})".format(n_fft, win_length))
    return np.pad(w, (lpad, n_fft - win_length - lpad))


def _cached(fn):
    """Memoise a table builder on its (hashable) arguments; the cached arrays are read-only.  Per-call API functions
    (`compute_fbank_feats`, `fbank`, `mfcc`, ...) rebuilt these tables on every call: 0.12 of 0.44 ms for one 10 s utterance."""
    @functools.lru_cache(maxsize=64)
    def build(*args):
        out = np.asarray(fn(*args))
        out.setflags(write=False)
        return out

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        if kw:
            return fn(*args, **kw)
        try:
            return build(*args)
        except TypeError:       # unhashable argument (an array-valued window, ...)
            return fn(*args)
    return wrapper


@_cached
def povey_window(n):
    """Symmetric hann ** 0.85 (``examples/conformer/dataset.py:126``)."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * math.pi * k / (n - 1))) ** 0.85


def _to_mel(f, kind):
    f = np.asarray(f, dtype=np.float64)
    if kind == "htk":
        return 2595.0 * np.log10(1.0 + f / 700.0)
    lin = f * 3.0 / 200.0
    with np.errstate(divide="ignore", invalid="ignore"):
        logpart = 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / math.log(6.4))
    return np.where(f >= 1000.0, logpart, lin)


def _from_mel(m, kind):
    m = np.asarray(m, dtype=np.float64)
    if kind == "htk":
        return 700.0 * (np.power(10.0, m / 2595.0) - 1.0)
    return np.where(m >= 15.0, 1000.0 * np.exp((m - 15.0) * (math.log(6.4) / 27.0)), m * 200.0 / 3.0)


@_cached
def hz_triangle_bank(n_stft, n_mels, sample_rate, f_min, f_max, norm=NormType.NONE, mel_type=MelType.HTK):
    """MelScale filterbank: triangles linear in Hz between mel-spaced corner
    frequencies; returned as ``[n_mels, n_stft]`` (rows = filters) for the C ABI."""
    norm, kind = NormType(norm).value, MelType(mel_type).value
    grid = np.linspace(0.0, float(sample_rate // 2), n_stft)
    corners = _from_mel(np.linspace(_to_mel(f_min, kind), _to_mel(f_max, kind), n_mels + 2), kind)
    lo, mid, hi = corners[:-2, None], corners[1:-1, None], corners[2:, None]
    rising = (grid[None, :] - lo) / (mid - lo)
    falling = (hi - grid[None, :]) / (hi - mid)
    bank = np.clip(np.minimum(rising, falling), 0.0, None)
    if norm == "slaney":
        bank = bank * (2.0 / (hi - lo))
    return bank


@_cached
def kaldi_triangle_bank(num_bins, n_fft, sample_rate, low_freq, high_freq):
    """Conformer-example filterbank (``examples/conformer/dataset.py:68-113``): triangles linear
    on the mel axis (1127 ln(1 + f/700)); Nyquist column zero; ``[num_bins, n_fft//2 + 1]``."""
    mel = lambda f: 1127.0 * np.log1p(np.asarray(f, dtype=np.float64) / 700.0)
    m_lo, m_hi = float(mel(low_freq)), float(mel(high_freq))
    step = (m_hi - m_lo) / (num_bins + 1)
    idx = np.arange(num_bins, dtype=np.float64)[:, None]
    left, centre, right = m_lo + idx * step, m_lo + (idx + 1.0) * step, m_lo + (idx + 2.0) * step
    bin_mel = mel(np.arange(n_fft // 2) * (float(sample_rate) / n_fft))[None, :]
    bank = np.clip(np.minimum((bin_mel - left) / (centre - left), (right - bin_mel) / (right - centre)), 0.0, None)
    return np.concatenate([bank, np.zeros((num_bins, 1))], axis=1)


@_cached
def dct_matrix(n_mfcc, n_mels, norm=NormMode.ORTHO):
    """``create_dct``: ``[n_mels, n_mfcc]`` float32."""
    norm = NormMode(norm).value
    pos = (np.arange(n_mels, dtype=np.float64) + 0.5)[:, None]
    order = np.arange(n_mfcc, dtype=np.float64)[None, :]
    mat = np.cos(math.pi / n_mels * pos * order)
    if norm == "ortho":
        mat[:, 0] *= math.sqrt(0.5)
        mat *= math.sqrt(2.0 / n_mels)
    else:
        mat *= 2.0
    return mat.astype(np.float32)
