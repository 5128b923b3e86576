"""float64 numpy restatement of the reference's front-end feature path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): the checker for the CUDA
path, never the thing shipped or measured (except as the labelled
``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``).

Every function cites the reference file:line (relative to the
mindspore-lab/mindaudio tree) whose behaviour it restates.  Code marked
[ms-op] restates an un-vendored ``mindspore==2.3.0`` C++ dataset op from its
API contract (torchaudio-lineage formula); everything else restates python
that lives in the reference repository and is checked against that python
by ``tests/test_oracle_vs_reference.py`` / the frozen goldens.
"""
from __future__ import annotations

import json
import math

import numpy as np
from scipy.signal import get_window as _scipy_get_window

# --------------------------------------------------------------------------
# windows / padding helpers
# --------------------------------------------------------------------------

_NP_PAD = {"constant": "constant", "edge": "edge", "reflect": "reflect", "symmetric": "symmetric"}


def periodic_window(name, win_length):
    """Periodic (``fftbins=True``) analysis window, float64.

    ``mindaudio/data/spectrum.py:173`` (``scipy.signal.get_window(window,
    win_length, fftbins=True)``); the [ms-op] Spectrogram uses the same periodic
    family (hann/hamming/blackman/bartlett/kaiser(beta=12)).
    """
    name = getattr(name, "value", name)
    if name == "kaiser":
        return _scipy_get_window(("kaiser", 12.0), win_length, fftbins=True)
    return _scipy_get_window(name, win_length, fftbins=True)


def pad_center(w, size):
    """Zero-pad ``w`` symmetrically to ``size`` -- ``spectrum.py:323-336``."""
    n = w.shape[-1]
    lpad = (size - n) // 2
    if lpad < 0:
        raise ValueError("Target size ({:d}) must be at least input size ({:d})".format(size, n))
    return np.pad(w, (lpad, size - n - lpad))


def frame_count(length, n_fft, hop, center):
    """Number of STFT columns -- follows ``spectrum.py:281-304`` on the padded signal."""
    eff = length + 2 * (n_fft // 2) if center else length
    return (eff - n_fft) // hop + 1


def _frames(x, n_fft, hop):
    """[.., L] -> [.., T, n_fft] strided view materialised (``spectrum.py:281-304``)."""
    t = (x.shape[-1] - n_fft) // hop + 1
    idx = np.arange(t)[:, None] * hop + np.arange(n_fft)[None, :]
    return x[..., idx]


# --------------------------------------------------------------------------
# a1  stft            mindaudio/data/spectrum.py:125-278
# --------------------------------------------------------------------------

def stft(waveforms, n_fft=512, win_length=None, hop_length=None, window="hann",
         center=True, pad_mode="constant", return_complex=True):
    """Naive form of ``spectrum.stft`` (``spectrum.py:125-278``).

    The reference's head/middle/tail evaluation (``:189-239``) is an optimisation
    of exactly this; its latent AttributeError at ``:237`` for some
    lengths with ``hop > n_fft/2`` is not reproduced -- the mathematically defined
    result is returned.  float64 math, complex64 result, ``[.., F, T]``.
    """
    x = np.asarray(waveforms)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    if hop_length < 1:
        raise ValueError("Invalid hop_length: {:d}".format(hop_length))
    w = pad_center(periodic_window(window, win_length), n_fft)
    if n_fft > x.shape[-1]:
        raise ValueError("n_fft={} is too large for input signal of length={}".format(n_fft, x.shape[-1]))
    x = x.astype(np.float64)
    if center:
        pw = [(0, 0)] * (x.ndim - 1) + [(n_fft // 2, n_fft // 2)]
        x = np.pad(x, pw, mode=pad_mode)
    fr = _frames(x, n_fft, hop_length) * w              # [.., T, n_fft]
    spec = np.fft.rfft(fr, axis=-1)                      # [.., T, F]
    spec = np.swapaxes(spec, -1, -2).astype(np.complex64)
    if return_complex:
        return spec
    return np.stack((spec.real, spec.imag), -1)


# --------------------------------------------------------------------------
# a2  istft           mindaudio/data/spectrum.py:346-494
# --------------------------------------------------------------------------

def window_sumsquare(window, n_frames, win_length, n_fft, hop_length):
    """``spectrum.py:477-494``."""
    n = n_fft + hop_length * (n_frames - 1)
    wsq = pad_center(periodic_window(window, win_length) ** 2, n_fft)
    out = np.zeros(n, dtype=np.float64)
    for t in range(n_frames):
        s = t * hop_length
        out[s:s + n_fft] += wsq[: max(0, min(n_fft, n - s))]
    return out


def istft(stft_matrix, n_fft=None, win_length=None, hop_length=None, window="hann",
          center=True, length=None):
    """``spectrum.py:346-474``: windowed irfft, overlap-add, /window-sum-square, trim.

    float64 arithmetic: the reference's environment (python 3.8/3.9, numpy 1.x --
    .github/workflows/ut_test.yaml:33-44) promotes complex64 to complex128 inside
    ``np.fft.irfft``; numpy >= 2 would transform complex64 in single precision, so the input
    is up-cast explicitly here."""
    z = np.asarray(stft_matrix).astype(np.complex128)
    if n_fft is None:
        n_fft = 2 * (z.shape[-2] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    w = pad_center(periodic_window(window, win_length), n_fft)
    if length:
        padded = length + int(n_fft) if center else length
        n_frames = min(z.shape[-1], int(np.ceil(padded / hop_length)))
    else:
        n_frames = z.shape[-1]
    n = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(z.shape[:-2] + (n,), dtype=np.float64)
    seg = np.fft.irfft(z[..., :n_frames], n=n_fft, axis=-2) * w[:, None]   # [.., n_fft, T]
    for t in range(n_frames):
        y[..., t * hop_length: t * hop_length + n_fft] += seg[..., t]
    wss = window_sumsquare(window, n_frames, win_length, n_fft, hop_length)
    nz = wss > 1e-9
    y[..., nz] /= wss[nz]
    if length is None:
        if center:
            y = y[..., n_fft // 2: -(n_fft // 2)]
    else:
        start = n_fft // 2 if center else 0
        y = y[..., start:]
        if y.shape[-1] > length:
            y = y[..., :length]
        elif y.shape[-1] < length:
            y = np.pad(y, [(0, 0)] * (y.ndim - 1) + [(0, length - y.shape[-1])])
    return y


# --------------------------------------------------------------------------
# a3  magphase        mindaudio/data/spectrum.py:701-735
# --------------------------------------------------------------------------

def magphase(z, power):
    """``spectrum.py:720-732`` (iscomplex=True branch); N-D accepted (superset)."""
    z = np.asarray(z)
    mag = np.abs(z)
    zero = mag == 0
    den = mag + zero
    phase = np.empty(z.shape, dtype=np.complex64)
    phase.real = z.real / den + zero
    phase.imag = z.imag / den
    mag = mag ** power
    return mag, phase


def magphase_real(x, power):
    """[ms-op] ``msaudio.Magphase(power)`` on ``[.., 2]`` (``spectrum.py:734-735``):
    magnitude = |z|**power, phase = atan2(im, re)."""
    x = np.asarray(x)
    mag = np.hypot(x[..., 0], x[..., 1]) ** power
    ang = np.arctan2(x[..., 1], x[..., 0])
    return mag, ang


# --------------------------------------------------------------------------
# a4/a5  amplitude_to_dB / dB_to_amplitude   spectrum.py:25-113
# --------------------------------------------------------------------------

def amplitude_to_dB(x, stype="power", ref=1.0, amin=1e-10, top_db=80.0):
    """``spectrum.py:59-90``.  Clamp group: 2-D -> the matrix; 3-D -> WHOLE batch
    (``channels = shape[-3]``, ``:81-86``); 4-D -> per leading item."""
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.complexfloating):
        raise UserWarning("amplitude_to_db was called on complex input")
    ref_value = ref(x) if callable(ref) else np.abs(ref)
    mult = 10.0 if stype == "power" else 20.0
    db = mult * np.log10(np.clip(x, a_min=amin, a_max=None))
    db = db - mult * np.log10(max(amin, ref_value))
    if top_db is not None:
        shape = db.shape
        channels = shape[-3] if len(shape) > 2 else 1
        g = db.reshape((-1, channels, shape[-2], shape[-1]))
        floor = np.amax(g, axis=(-3, -2, -1)) - top_db
        db = np.maximum(g, floor.reshape((-1, 1, 1, 1))).reshape(shape)
    return db


def dB_to_amplitude(x, ref, power):
    """``spectrum.py:108-113``."""
    ref_value = ref(x) if callable(ref) else np.abs(ref)
    return ref_value * np.power(np.power(10.0, 0.1 * np.asarray(x)), power)


# --------------------------------------------------------------------------
# a6  spectrogram [ms-op]   call sites spectrum.py:594-606, 673-684
# --------------------------------------------------------------------------

def spectrogram(waveforms, n_fft=400, win_length=None, hop_length=None, pad=0, window="hann",
                power=2.0, normalized=False, center=True, pad_mode="reflect", onesided=True):
    """[ms-op] ``msaudio.Spectrogram`` = ``torchaudio.functional.spectrogram`` semantics
     inside the reference wrapper's defaults (``spectrum.py:590-591``:
    ``win = n_fft``, ``hop = win // 2``).  Output dtype follows the input
    (float64 stays float64, else float32)."""
    x = np.asarray(waveforms)
    out_dtype = np.float64 if x.dtype == np.float64 else np.float32
    win_length = win_length if win_length else n_fft
    hop_length = hop_length if hop_length else win_length // 2
    pad_mode = _NP_PAD[getattr(pad_mode, "value", pad_mode)]
    x = x.astype(np.float64)
    lead = [(0, 0)] * (x.ndim - 1)
    if pad > 0:
        x = np.pad(x, lead + [(pad, pad)])
    w = pad_center(periodic_window(window, win_length), n_fft)
    if center:
        x = np.pad(x, lead + [(n_fft // 2, n_fft // 2)], mode=pad_mode)
    fr = _frames(x, n_fft, hop_length) * w
    spec = np.fft.rfft(fr, axis=-1) if onesided else np.fft.fft(fr, axis=-1)
    if normalized:
        spec = spec / np.sqrt(np.sum(w ** 2))
    mag = np.abs(spec)
    if power == 2.0:
        out = spec.real ** 2 + spec.imag ** 2
    elif power == 1.0:
        out = mag
    else:
        out = mag ** power
    return np.swapaxes(out, -1, -2).astype(out_dtype)


# --------------------------------------------------------------------------
# a7  melscale / melspectrogram [ms-op]   call sites spectrum.py:686-698, 771-774
# --------------------------------------------------------------------------

def _hz_to_mel(f, mel_type):
    f = np.asarray(f, dtype=np.float64)
    if mel_type == "htk":
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m, mel_type):
    m = np.asarray(m, dtype=np.float64)
    if mel_type == "htk":
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def melscale_fbanks(n_stft, f_min, f_max, n_mels, sample_rate, norm="none", mel_type="htk"):
    """[ms-op] filterbank of ``msaudio.MelScale`` = ``torchaudio.functional.melscale_fbanks``
    (SURVEY.md A6): triangles linear in Hz, ``[n_stft, n_mels]``, float64."""
    norm = getattr(norm, "value", norm)
    mel_type = getattr(mel_type, "value", mel_type)
    all_freqs = np.linspace(0, sample_rate // 2, n_stft)
    m_pts = np.linspace(_hz_to_mel(f_min, mel_type), _hz_to_mel(f_max, mel_type), n_mels + 2)
    f_pts = _mel_to_hz(m_pts, mel_type)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return fb


def melscale(spec, n_mels=128, sample_rate=16000, f_min=0, f_max=None, n_stft=201,
             norm="none", mel_type="htk"):
    """[ms-op] ``msaudio.MelScale`` applied as ``(spec^T @ fb)^T`` (``spectrum.py:770-774``)."""
    spec = np.asarray(spec)
    out_dtype = np.float64 if spec.dtype == np.float64 else np.float32
    f_max = f_max if f_max is not None else sample_rate // 2
    fb = melscale_fbanks(n_stft, f_min, f_max, n_mels, sample_rate, norm, mel_type)
    mel = np.swapaxes(np.swapaxes(spec.astype(np.float64), -1, -2) @ fb, -1, -2)
    return mel.astype(out_dtype)


def melspectrogram(waveforms, n_fft=400, win_length=None, hop_length=None, pad=0, window="hann",
                   power=2.0, normalized=False, center=True, pad_mode="reflect", onesided=True,
                   n_mels=128, sample_rate=16000, f_min=0, f_max=None, norm="none", mel_type="htk"):
    """``spectrum.py:609-698`` around the two [ms-op]s; ``hop = win // 2`` (``:665-666``)."""
    win_length = win_length if win_length is not None else n_fft
    hop_length = hop_length if hop_length is not None else win_length // 2
    spec = spectrogram(waveforms, n_fft, win_length, hop_length, pad, window, power,
                       normalized, center, pad_mode, onesided)
    return melscale(spec, n_mels, sample_rate, f_min, f_max, n_fft // 2 + 1, norm, mel_type)


# --------------------------------------------------------------------------
# a8/a9, f1  fbank / mfcc / deltas / context     features.py:69-373
# --------------------------------------------------------------------------

def create_dct(n_mfcc, n_mels, norm="ortho"):
    """[ms-op] ``mindspore.dataset.audio.utils.create_dct`` = ``torchaudio.functional.create_dct``
    (SURVEY.md A7): ``[n_mels, n_mfcc]``; returned in float64 here (the op returns float32)."""
    norm = getattr(norm, "value", norm)
    n = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    dct = np.cos(math.pi / n_mels * (n + 0.5) * k)
    if norm in (None, "none"):
        dct = dct * 2.0
    else:
        dct[0] *= 1.0 / math.sqrt(2.0)
        dct = dct * math.sqrt(2.0 / n_mels)
    return dct.T.copy()


def compute_deltas(specgram, win_length=5, pad_mode="edge"):
    """[ms-op] ``msaudio.ComputeDeltas`` = ``torchaudio.functional.compute_deltas`` (A8);
    call site ``features.py:191-193``."""
    x = np.asarray(specgram)
    out_dtype = np.float64 if x.dtype == np.float64 else np.float32
    x = x.astype(np.float64)
    n = (win_length - 1) // 2
    denom = n * (n + 1) * (2 * n + 1) / 3.0
    mode = _NP_PAD[getattr(pad_mode, "value", pad_mode)]
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(n, n)], mode=mode)
    t = x.shape[-1]
    out = np.zeros_like(x)
    for k in range(-n, n + 1):
        out += k * xp[..., n + k: n + k + t]
    return (out / denom).astype(out_dtype)


def context_window(x, left_frames=0, right_frames=0):
    """``features.py:94-155`` (grouped identity-kernel Conv1d) restated as a gather (A8):
    ``out[f*C + c, t] = x[f, t + c + roll - max(l, r)]`` with zero padding, float32."""
    x = np.asarray(x)
    if x.ndim not in (2, 3, 4):
        raise TypeError("Input dimension must be 2, 3 or 4, but got {}".format(x.ndim))
    l, r = left_frames, right_frames
    csize = l + r + 1
    mf = max(l, r)
    ksz = 2 * mf + 1
    shift = r - l
    # tap position of output channel c inside the kernel (np.eye rolled right by shift if > 0)
    taps = [(c + (shift if shift > 0 else 0)) % ksz for c in range(csize)]

    def one(m):  # m: [F, T] -> [F*C, T]
        f, t = m.shape
        mp = np.pad(m.astype(np.float32), ((0, 0), (mf, mf)))
        out = np.empty((f, csize, t), dtype=np.float32)
        for c, tap in enumerate(taps):
            out[:, c, :] = mp[:, tap: tap + t]
        return out.reshape(f * csize, t)

    if x.ndim == 2:
        # the reference never squeezes the batch axis it added: `len(x_shape) == 2` is tested on the
        # EXPANDED shape (features.py:104-106, 153-154), so 2-D input comes back as [1, F*C, T]
        return one(x)[None]
    if x.ndim == 3:
        return np.stack([one(m) for m in x])
    # 4-D: [B, C, F, T]: the reference moves C last, folds (B, F) and convolves over T per C
    b, ch, f, t = x.shape
    xt = x.transpose((0, 2, 3, 1))                       # [B, F, T, C]
    # reference reshapes [B,F,T,C] -> (B*T, F, C) and convolves along C (features.py:125-126);
    # restated literally:
    xr = xt.reshape((b * t, f, ch))
    ctx = np.stack([one(m) for m in xr])                 # [B*T, F*csize, C]
    ctx = ctx.reshape((b, ctx.shape[1], t, ctx.shape[-1]))
    return ctx.transpose((0, 3, 1, 2))


def fbank(waveforms, deltas=False, context=False, n_mels=40, n_fft=400, sample_rate=16000,
          f_min=0.0, f_max=None, left_frames=5, right_frames=5, win_length=None,
          hop_length=None, window="hann"):
    """``features.py:196-270``."""
    mel = melspectrogram(waveforms, n_fft=n_fft, win_length=win_length, hop_length=hop_length,
                         window=window, n_mels=n_mels, sample_rate=sample_rate, f_min=f_min, f_max=f_max)
    fb = amplitude_to_dB(mel, stype="power", ref=1.0, top_db=80.0)
    if deltas:
        d1 = compute_deltas(fb)
        d2 = compute_deltas(d1)
        fb = np.concatenate((fb, d1, d2), axis=-2)
    if context:
        fb = context_window(fb, left_frames, right_frames)
    return fb


def mfcc(waveforms, deltas=True, context=True, n_mels=23, n_mfcc=20, n_fft=400, sample_rate=16000,
         f_min=0.0, f_max=None, left_frames=5, right_frames=5, win_length=None, hop_length=None,
         norm="ortho", log_mels=False):
    """``features.py:273-373``."""
    if n_mfcc > n_mels:
        raise ValueError("The number of MFCC coefficients must be no more than # mel bins.")
    dct = create_dct(n_mfcc, n_mels, norm).astype(np.float32)     # the op returns float32
    mel = melspectrogram(waveforms, sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, f_min=f_min,
                         f_max=f_max, win_length=win_length, hop_length=hop_length)
    if log_mels:
        mel = np.log(mel + 1e-6)
    else:
        mel = amplitude_to_dB(mel, stype="power", ref=1.0, top_db=80.0)
    if mel.ndim not in (2, 3, 4):
        raise TypeError("Unsupported MelSpectrogram shape {}".format(mel.ndim))
    out = np.swapaxes(np.matmul(np.swapaxes(mel, -1, -2), dct), -1, -2)
    if deltas:
        d1 = compute_deltas(out)
        d2 = compute_deltas(d1)
        out = np.concatenate((out, d1, d2), axis=-2)
    if context:
        out = context_window(out, left_frames, right_frames)
    return out


# --------------------------------------------------------------------------
# a10  conformer Kaldi-like fbank    examples/conformer/dataset.py:56-168
# --------------------------------------------------------------------------

def kaldi_mel(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def kaldi_mel_banks(num_bins=80, n_fft=512, sample_rate=16000.0, low=20.0, high=8000.0):
    """``examples/conformer/dataset.py:68-113``: triangles linear on the MEL axis,
    ``[num_bins, n_fft//2 + 1]`` with the last column zero (``:111``)."""
    nb = n_fft // 2
    width = sample_rate / n_fft
    lo, hi = float(kaldi_mel(low)), float(kaldi_mel(high))
    delta = (hi - lo) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.float64)[:, None]
    left, centre, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
    mel = kaldi_mel(width * np.arange(nb))[None, :]
    up = (mel - left) / (centre - left)
    down = (right - mel) / (right - centre)
    w = np.maximum(0.0, np.where(up > down, down, up))
    return np.pad(w, ((0, 0), (0, 1)))


def povey_window(n=400):
    """``examples/conformer/dataset.py:126``: ``np.hanning(n) ** 0.85`` (symmetric)."""
    return np.power(np.hanning(n), 0.85)


def conformer_fbank(wav, sample_rate=16000, frame_len_ms=25, frame_shift_ms=10, mel_bin=80,
                    n_fft=512, dither=0.0, seed=0, utt_id=0):
    """``examples/conformer/dataset.py:159-168`` (+ ``:117-156``), float64, ``[T, mel_bin]``.

    pre-emphasis over the whole signal -> povey frames -> ONE scalar mean over all
    frame entries removed (``:165``) -> |rfft(., 512)|^2 -> mel -> 0 -> eps -> ln.
    ``dither`` is OUR extension (reference raises NotImplementedError,
    ``dataset.py:557-558``); 0 leaves the reference path untouched.
    """
    x = np.asarray(wav, dtype=np.float64)
    if dither:
        x = x + dither * dither_noise(x.shape[-1], seed, utt_id).astype(np.float64)
    y = np.append(x[0], x[1:] - 0.97 * x[:-1])
    flen = sample_rate * frame_len_ms // 1000
    fshift = sample_rate * frame_shift_ms // 1000
    t = int(np.floor((y.size - flen) / fshift) + 1)
    if t <= 0:
        return np.zeros((0, mel_bin))
    idx = np.arange(t)[:, None] * fshift + np.arange(flen)[None, :]
    fr = y[idx] * povey_window(flen)
    fr = fr - np.mean(fr)
    spec = np.abs(np.fft.rfft(fr, n=n_fft)) ** 2
    bank = kaldi_mel_banks(mel_bin, n_fft, float(sample_rate), 20.0, 8000.0)
    feats = spec @ bank.T
    feats = np.where(feats == 0, np.finfo(float).eps, feats)
    return np.log(feats)


# --------------------------------------------------------------------------
# a11-a13  CMVN family
# --------------------------------------------------------------------------

def utt_cmvn(x, mean_norm=True, std_norm=True):
    """``examples/ECAPA-TDNN/spec_augment.py:43-70`` (norm_type='sentence'):
    per utterance ``[T, D]``, per feature dim, population std, NO eps."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(axis=0) if mean_norm else 0.0
    std = x.std(axis=0) if std_norm else 1.0
    return (x - mean) / std


def scalar_norm(mag):
    """``examples/deepspeech2/dataset.py:43-47``: ``log1p`` then scalar mean/std over the matrix."""
    m = np.log1p(np.asarray(mag, dtype=np.float64))
    return (m - m.mean()) / m.std()


def cmvn_stats(feats_list):
    """``examples/conformer/compute_cmvn_stats.py:61-63, 104-112``: (N, sum x, sum x^2), float64."""
    d = feats_list[0].shape[1]
    s1, s2, n = np.zeros(d), np.zeros(d), 0
    for f in feats_list:
        f = np.asarray(f, dtype=np.float64)
        s1 += f.sum(axis=0)
        s2 += np.square(f).sum(axis=0)
        n += f.shape[0]
    return n, s1, s2


def cmvn_stats_json(n, s1, s2):
    """On-disk format of ``compute_cmvn_stats.py:121-128``."""
    return json.dumps({"mean_stat": list(np.asarray(s1).tolist()),
                       "var_stat": list(np.asarray(s2).tolist()), "frame_num": int(n)})


def cmvn_from_stats(n, s1, s2):
    """``mindaudio/utils/load_files.py:19-28``: mean, istd with the 1e-20 variance floor."""
    mean = np.asarray(s1, dtype=np.float64) / n
    var = np.asarray(s2, dtype=np.float64) / n - mean * mean
    var = np.where(var < 1.0e-20, 1.0e-20, var)
    return mean, 1.0 / np.sqrt(var)


def global_cmvn_apply(x, mean, istd, norm_var=True):
    """``mindaudio/models/layers/cmvn.py:33-36`` in float32 (``asr_model.py:304-305``)."""
    x = np.asarray(x, dtype=np.float32) - np.asarray(mean, dtype=np.float32)
    if norm_var:
        x = x * np.asarray(istd, dtype=np.float32)
    return x


# --------------------------------------------------------------------------
# A11  dither (ours; the reference has none)
# --------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011), vectorised over the counter arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dither_noise(n, seed, utt_id):
    """Standard-normal dither g[0..n): counter = (i, 0, utt_id, 0), key = (seed_lo, seed_hi);
    Box-Muller on the top 24 bits of the first two words (exact in float32 and float64):
    u1 = ((w0 >> 8) + 1) * 2^-24 in (0, 1], u2 = (w1 >> 8) * 2^-24,
    g = sqrt(-2 ln u1) * cos(2 pi u2).  float32 result (float64 math here)."""
    i = np.arange(n, dtype=np.uint64)
    w0, w1, _, _ = philox4x32_10(i.astype(np.uint32), (i >> np.uint64(32)).astype(np.uint32),
                                 np.uint32(utt_id & 0xFFFFFFFF), np.uint32(0),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = ((w0 >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (w1 >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# collate (scope row f3): mindaudio/utils/common.py:10-52, mindaudio/utils/mask.py:44-67
# ---------------------------------------------------------------------------------------------
def pad_sequence(sequences, batch_first=True, padding_value=0, padding_max_len=None, atype=np.int32):
    trailing = tuple(sequences[0].shape[1:])
    max_len = padding_max_len if padding_max_len is not None else max(s.shape[0] for s in sequences)
    dims = ((len(sequences), max_len) if batch_first else (max_len, len(sequences))) + trailing
    out = np.full(dims, fill_value=padding_value).astype(atype)
    for i, seq in enumerate(sequences):
        n = min(seq.shape[0], max_len)
        if batch_first:
            out[i, :n] = seq[:n]
        else:
            out[:n, i] = seq[:n]
    return out


def make_pad_mask(lengths, max_len=0):
    lengths = np.asarray(lengths)
    max_len = max_len if max_len > 0 else int(lengths.max())
    return np.arange(max_len)[None, :] >= lengths[:, None]


def conformer_collate_x(feats, max_src_len):
    """xs_pad, xs_lengths, xs_masks of examples/conformer/dataset.py:563-569, 616-621."""
    xs_pad = pad_sequence(feats, batch_first=True, padding_value=0.0, padding_max_len=max_src_len, atype=np.float32)
    xs_lengths = np.array([x.shape[0] for x in feats], dtype=np.int32)
    xs_masks = np.expand_dims(~make_pad_mask(xs_lengths, max_len=max_src_len), 1).astype(np.float32)
    return xs_pad, xs_lengths, xs_masks


# ---------------------------------------------------------------------------------------------
# feature-domain ops of the pipelines (scope row f4)
# ---------------------------------------------------------------------------------------------
def sliding_window_cmn(x, cmn_window=600, min_cmn_window=100, center=False, norm_vars=False):
    """[ms-op] ``msaudio.SlidingWindowCmn`` (processing.py:380-407) = Kaldi's sliding-window CMN as written in
    torchaudio.functional.sliding_window_cmn: window rule per frame, mean (and variance) over the window.  Direct
    window sums in float64 (torchaudio updates them incrementally in the input dtype)."""
    x = np.asarray(x)
    out_dtype = np.float64 if x.dtype == np.float64 else np.float32
    xf = x.astype(np.float64)
    T = xf.shape[-2]
    out = np.zeros_like(xf)
    for t in range(T):
        if center:
            ws = t - cmn_window // 2
            we = ws + cmn_window
        else:
            ws, we = t - cmn_window, t + 1
        if ws < 0:
            we -= ws
            ws = 0
        if not center and we > t:
            we = max(t + 1, min_cmn_window)
        if we > T:
            ws -= we - T
            we = T
            ws = max(ws, 0)
        win = xf[..., ws:we, :]
        n = we - ws
        mean = win.sum(axis=-2) / n
        y = xf[..., t, :] - mean
        if norm_vars:
            if n == 1:
                y = np.zeros_like(y)
            else:
                var = (win ** 2).sum(axis=-2) / n - mean ** 2
                y = y * var ** -0.5
        out[..., t, :] = y
    return out.astype(out_dtype)


def spectral_centroid(waveforms, sample_rate, n_fft=400, win_length=None, hop_length=None, pad=0, window="hann"):
    """[ms-op] ``msaudio.SpectralCentroid`` (features.py:22-66) = torchaudio.functional.spectral_centroid."""
    x = np.asarray(waveforms)
    out_dtype = np.float64 if x.dtype == np.float64 else np.float32
    win_length = win_length if win_length else n_fft
    hop_length = hop_length if hop_length else win_length // 2
    spec = spectrogram(x.astype(np.float64), n_fft, win_length, hop_length, pad, window, 1.0, False, True, "reflect")
    freqs = np.linspace(0, sample_rate // 2, 1 + n_fft // 2).reshape((-1, 1))
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((freqs * spec).sum(axis=-2) / spec.sum(axis=-2)).astype(out_dtype)


def spec_aug(xs, spec_aug_conf, rng):
    """examples/conformer/dataset.py:493-534, in place; ``rng`` is a ``random.Random`` (the reference uses the module)."""
    num_t_mask = spec_aug_conf.get("num_t_mask", 0)
    num_f_mask = spec_aug_conf.get("num_f_mask", 0)
    max_t = spec_aug_conf.get("max_t", 0)
    max_f = spec_aug_conf.get("max_f", 0)
    for x in xs:
        max_frames, max_freq = x.shape[0], x.shape[1]
        for _ in range(num_t_mask):
            start = rng.randint(0, max_frames - 1)
            length = rng.randint(1, max_t)
            end = min(max_frames, start + length)
            if rng.randint(1, 100) > 20:
                x[start:end, :] = 0
        for _ in range(num_f_mask):
            start = rng.randint(0, max_freq - 1)
            length = rng.randint(1, max_f)
            end = min(max_freq, start + length)
            if rng.randint(1, 100) > 20:
                x[:, start:end] = 0
    return xs


def mask_along_axis(spec, mask_param, mask_start, mask_value, axis):
    """[ms-op, recalled] deterministic branch of ``msaudio.FrequencyMasking`` / ``TimeMasking`` (iid_masks=False):
    ``mask_param`` consecutive indices from ``mask_start`` along ``axis`` (-2 = frequency, -1 = time) set to value."""
    out = np.array(spec, copy=True)
    idx = [slice(None)] * out.ndim
    idx[axis] = slice(mask_start, mask_start + mask_param)
    out[tuple(idx)] = mask_value
    return out


# ---------------------------------------------------------------------------------------------
# phase vocoder / time_stretch (scope row f2): mindaudio/data/augment.py:795-871
# ---------------------------------------------------------------------------------------------
def phase_vocoder(matrix, rate, hop_length=None, n_fft=None):
    """augment.py:828-871 with its dtype flow: complex64 in -> float32 phase accumulator updated in float64."""
    matrix = np.asarray(matrix)
    if n_fft is None:
        n_fft = 2 * (matrix.shape[-2] - 1)
    if hop_length is None:
        hop_length = int(n_fft // 4)
    time_steps = np.arange(0, matrix.shape[-1], rate, dtype=np.float64)
    shape = list(matrix.shape)
    shape[-1] = len(time_steps)
    d_stretch = np.zeros_like(matrix, shape=shape)
    phi_advance = np.linspace(0, np.pi * hop_length, matrix.shape[-2])
    phase_acc = np.angle(matrix[..., 0])
    padding = [(0, 0) for _ in matrix.shape]
    padding[-1] = (0, 2)
    matrix = np.pad(matrix, padding, mode="constant")
    for t, step in enumerate(time_steps):
        columns = matrix[..., int(step): int(step + 2)]
        alpha = np.mod(step, 1.0)
        mag = (1.0 - alpha) * np.abs(columns[..., 0]) + alpha * np.abs(columns[..., 1])
        d_stretch[..., t] = (np.cos(phase_acc) + 1j * np.sin(phase_acc)) * mag
        dphase = np.angle(columns[..., 1]) - np.angle(columns[..., 0]) - phi_advance
        dphase = dphase - 2.0 * np.pi * np.round(dphase / (2.0 * np.pi))
        phase_acc += phi_advance + dphase
    return d_stretch


def time_stretch(waveforms, rate):
    if rate <= 0:
        raise ValueError("rate must be a positive number")
    waveforms = np.asarray(waveforms)
    spec = stft(waveforms)
    return istft(phase_vocoder(spec, rate), length=int(round(waveforms.shape[-1] / rate)))


# ---------------------------------------------------------------------------------------------
# harmonic / percussive separation (scope row f2): mindaudio/data/features.py:438-559
# ---------------------------------------------------------------------------------------------
def median_filter_1d(x, size, axis):
    """scipy.ndimage.median_filter(x, size=[.., size at axis ..], mode="reflect") restated with numpy: window
    [i - size//2, i - size//2 + size), rank size//2, symmetric (half-sample) boundary."""
    x = np.asarray(x)
    lo = size // 2
    pad = [(0, 0)] * x.ndim
    pad[axis] = (lo, size - 1 - lo)
    xp = np.pad(x, pad, mode="symmetric")
    win = np.lib.stride_tricks.sliding_window_view(xp, size, axis=axis)
    return np.sort(win, axis=-1)[..., size // 2]


def soft_mask(x_input, x_ref, power=1, split_zeros=False):
    input_type = x_input.dtype if np.issubdtype(x_input.dtype, np.floating) else np.float32
    z = np.maximum(x_input, x_ref).astype(input_type)
    bad_idx = z < np.finfo(input_type).tiny
    z[bad_idx] = 1
    if not np.isfinite(power):
        return x_input > x_ref
    mask = (x_input / z) ** power
    ref_mask = (x_ref / z) ** power
    good_idx = ~bad_idx
    mask[good_idx] /= mask[good_idx] + ref_mask[good_idx]
    mask[bad_idx] = 0.5 if split_zeros else 0.0
    return mask


def hpss(spectrogram, kernel_size=31, power=2.0, mask=False, margin=1.0):
    spectrogram = np.asarray(spectrogram)
    if np.iscomplexobj(spectrogram):
        spectrogram, phase = magphase(spectrogram, power=1)
    else:
        phase = 1
    margin_h, margin_p = (margin[0], margin[1]) if not np.isscalar(margin) else (margin, margin)
    win_h, win_p = (kernel_size[0], kernel_size[1]) if not np.isscalar(kernel_size) else (kernel_size, kernel_size)
    harm = median_filter_1d(spectrogram, win_h, -1)
    perc = median_filter_1d(spectrogram, win_p, -2)
    split_zeros = margin_h == 1 and margin_p == 1
    mask_h = soft_mask(harm, perc * margin_h, power=power, split_zeros=split_zeros)
    mask_p = soft_mask(perc, harm * margin_p, power=power, split_zeros=split_zeros)
    if mask:
        return mask_h, mask_p
    return (spectrogram * mask_h) * phase, (spectrogram * mask_p) * phase


def harmonic(y_input, **kwargs):
    y_input = np.asarray(y_input)
    return istft(hpss(stft(y_input, n_fft=2048, pad_mode="constant"), **kwargs)[0], length=y_input.shape[-1])



# --------------------------------------------------------------------------
# WAV decode ("next" row f3): mindaudio/data/io.py:347-747
# --------------------------------------------------------------------------

class WavWarning(UserWarning):
    """stands for io.py:339 WavFileWarning"""


def wav_read(raw, offset=0.0, duration=None, filelike=False):
    """``mindaudio.data.io.read`` over the bytes of a file (io.py:552-747), returning ``(audio, samplerate, notes)``.

    ``filelike`` selects the branch the reference takes for objects without a file descriptor (``np.fromfile`` raises
    ``io.UnsupportedOperation`` -> ``read(size)``, io.py:500-503).  ``notes`` lists the warnings the reference would
    emit ("unknown", "eof", "incomplete")."""
    import struct
    raw = bytes(raw)
    pos = 0
    notes = []

    def take(k):                      # file.read(k)
        nonlocal pos
        out = raw[pos:pos + k] if pos < len(raw) else b""
        pos += len(out)
        return out

    magic = take(4)                   # io.py:652-665
    if magic == b"RIFF":
        e = "<"
    elif magic == b"RIFX":
        e = ">"
    else:
        raise ValueError("File format %r not understood. Only 'RIFF' and 'RIFX' supported." % (magic,))
    total = struct.unpack(e + "I", take(4))[0] + 8            # io.py:668-670
    if take(4) != b"WAVE":
        raise TypeError("exceptions must derive from BaseException")   # io.py:676 raises a str

    def skip_chunk():                 # io.py:520-538
        nonlocal pos
        d = take(4)
        if d:
            size = struct.unpack(e + "I", d)[0]
            pos += size + (size & 1)

    fmt_seen = data_seen = False
    audio = None
    tag = channels = rate = align = depth = None
    while pos < total:                # io.py:680-736
        cid = take(4)
        if not cid:
            if data_seen:
                notes.append("eof")
                break
            raise ValueError("Unexpected end of file.")
        if len(cid) < 4:
            if fmt_seen and data_seen:
                notes.append("incomplete")
            else:
                raise ValueError("Incomplete chunk ID: %r" % (cid,))
        if cid == b"fmt ":            # io.py:347-424
            fmt_seen = True
            csize = struct.unpack(e + "I", take(4))[0]
            if csize < 16:
                raise ValueError("Binary structure of wave file is not compliant")
            tag, channels, rate, bps_sec, align, depth = struct.unpack(e + "HHIIHH", take(16))
            used = 16
            if tag == 0xFFFE and csize >= used + 2:
                ext = struct.unpack(e + "H", take(2))[0]
                used += 2
                if ext < 22:
                    raise ValueError("Binary structure of wave file is not compliant")
                blob = take(22)
                used += 22
                guid = blob[6:22]
                tail = (b"\x00\x00\x00\x10" if e == ">" else b"\x00\x00\x10\x00") + b"\x80\x00\x00\xAA\x00\x38\x9B\x71"
                if guid.endswith(tail):
                    tag = struct.unpack(e + "I", guid[:4])[0]
            if tag not in (1, 3):
                raise ValueError("Unknown wave file format: %#06x" % tag)
            if csize > used:
                take(csize - used)
            pos += csize & 1
            if tag == 1 and bps_sec != rate * align:
                raise ValueError("WAV header is invalid: nAvgBytesPerSec must equal product of nSamplesPerSec and nBlockAlign")
        elif cid == b"data":          # io.py:427-517
            data_seen = True
            if not fmt_seen:
                raise ValueError("No fmt chunk before data")
            size = struct.unpack(e + "I", take(4))[0]
            width = align // channels
            n_samples = size // width
            if tag == 1:
                if 1 <= depth <= 8:
                    dt = "u1"
                elif width in (3, 5, 6, 7):
                    dt = "V1"
                elif depth <= 64:
                    dt = e + "i%d" % width
                else:
                    raise ValueError("Unsupported bit depth: the WAV file has %d-bit integer data." % depth)
            else:
                if depth not in (32, 64):
                    raise ValueError("Unsupported bit depth: the WAV file has %d-bit floating-point data." % depth)
                dt = e + "f%d" % width
            skip = 0
            if offset > 0:
                skip = int(offset * rate)
                take(skip)                                    # BYTES, not samples (io.py:489-491)
            start = pos
            item = np.dtype(dt).itemsize
            if not filelike:
                count = size if dt == "V1" else n_samples
                if skip <= count:
                    count -= skip
                if duration and duration * rate < count:
                    count = int(duration * rate)
                got = min(count, max(0, len(raw) - start) // item)
                data = np.frombuffer(raw, dtype=dt, count=got, offset=start) if got else np.zeros(0, dtype=dt)
                pos = start + got * item
            else:
                data = np.frombuffer(take(size), dtype=dt)
            if dt == "V1":                                    # io.py:505-512
                wide = e + ("i4" if width == 3 else "i8")
                nb = np.dtype(wide).itemsize
                box = np.zeros((len(data) // width, nb), dtype="V1")
                if e == ">":
                    box[:, :width] = data.reshape((-1, width))
                else:
                    box[:, nb - width:] = data.reshape((-1, width))
                data = box.view(wide).reshape(box.shape[:-1])
            pos += size & 1
            audio = np.array(data)
            if channels > 1:
                audio = audio.reshape(-1, channels)
        elif cid in (b"fact", b"LIST", b"JUNK", b"Fake"):
            skip_chunk()
        else:
            notes.append("unknown")
            skip_chunk()
    if audio is None:
        raise UnboundLocalError("cannot access local variable 'audio' where it is not associated with a value")
    if audio.dtype == "int32":        # io.py:741-746
        audio = audio / 2147483648
    elif audio.dtype == "int16":
        audio = audio / 32768
    return audio, rate, notes


# --------------------------------------------------------------------------
# pitch shift ("next" row f2): mindaudio/data/augment.py:874-901, processing.py:132-186
# --------------------------------------------------------------------------

def resample(waveform, orig_freq=16000, new_freq=16000, res_type="fft"):
    """``processing.py:132-186``, "fft" / "scipy" branch: the reference hands the work to ``scipy.signal.resample``
    (third-party, ``requirements.txt``), and so does the oracle -- same call, same arguments."""
    import scipy.signal
    if orig_freq == new_freq:
        return waveform
    assert res_type in ("scipy", "fft")
    ratio = float(new_freq) / orig_freq
    n_samples = int(np.ceil(waveform.shape[-1] * ratio))
    return np.asarray(scipy.signal.resample(waveform, n_samples, axis=-1), dtype=waveform.dtype)


def pad_shape(y_shift, data_shape):
    """``spectrum.py:307-320``: crop or zero-pad the last axis to ``data_shape`` samples."""
    n = y_shift.shape[-1]
    if n > data_shape:
        return y_shift[..., :data_shape]
    if n < data_shape:
        return np.pad(y_shift, [(0, 0)] * (y_shift.ndim - 1) + [(0, data_shape - n)], mode="constant")
    return y_shift


def pitch_shift(waveforms, sr, n_steps, bins_per_octave=12):
    """``augment.py:874-901``."""
    rate = 2.0 ** (-float(n_steps) / bins_per_octave)
    stretched = time_stretch(waveforms, rate)
    y = resample(stretched, orig_freq=float(sr) / rate, new_freq=sr)
    return pad_shape(y, stretched.shape[-1])
