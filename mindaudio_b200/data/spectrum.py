"""Drop-in for ``mindaudio.data.spectrum`` (mindaudio/data/spectrum.py) on B200.

Same names, argument order, defaults, numpy-in / numpy-out semantics and error behaviour as the
reference; the arithmetic runs in libmafe.so (hand-written sm_100a kernels) through ctypes.
There is no CPU fallback: without the library / a B200 these functions raise.

Reference quirks kept on purpose: ``stft`` defaults to ``pad_mode="constant"``
and ``hop = win // 4`` while the Spectrogram wrappers default to reflect and ``win // 2``; ``stft``
returns complex64 whatever the input dtype; ``amplitude_to_dB`` on a 3-D batch clamps against the
max of the WHOLE batch; it raises ``UserWarning`` on complex input.  Fixed: the AttributeError of
``spectrum.py:237`` (some lengths with ``hop > n_fft/2``) -- the defined result is returned.

The small host utilities ``_pad_center``, ``_pad_shape``, ``compute_amplitude`` and ``resynthesize`` (a few lines each, off the hot
path) follow the reference's bodies closely on purpose: other reference modules import them from here and rely on their exact
shapes, dtypes and error messages.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L
from .. import _tables as T
from .._engine import get_engine
from .._enums import BorderType, MelType, NormType, WindowType

__all__ = [
    "amplitude_to_dB",
    "dB_to_amplitude",
    "stft",
    "istft",
    "compute_amplitude",
    "spectrogram",
    "melspectrogram",
    "magphase",
    "melscale",
    "resynthesize",
]

MAX_MEM_BLOCK = 2**8 * 2**10


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _out_dtype(x):
    return np.float64 if np.asarray(x).dtype == np.float64 else np.float32


def _flatten_batch(x):
    """[..., T] -> ([B, T] float32 C-contiguous, leading shape)."""
    x = np.asarray(x)
    lead = x.shape[:-1]
    return np.ascontiguousarray(x.reshape((-1, x.shape[-1])), dtype=np.float32), lead


def _dense_offsets(b, length):
    return np.arange(b + 1, dtype=np.int64) * length


# ------------------------------------------------------------------------------------------
def amplitude_to_dB(wavform, stype="power", ref=1.0, amin=1e-10, top_db=80.0):
    """``spectrum.py:25-90``.  Group of the ``top_db`` clamp: last three dims with
    ``channels = shape[-3]`` if ndim > 2 else 1 (``:81-86``)."""
    wavform = np.asarray(wavform)
    if np.issubdtype(wavform.dtype, np.complexfloating):
        raise UserWarning(
            "amplitude_to_db was called on complex input so phase "
            "information will be discarded. To suppress this warning, "
            "call amplitude_to_db(np.abs(D)**2) instead."
        )
    ref_value = ref(wavform) if callable(ref) else np.abs(ref)
    multiplier = 10.0 if stype == "power" else 20.0
    offset = multiplier * np.log10(max(amin, ref_value))
    shape = wavform.shape
    if top_db is not None:
        channels = shape[-3] if len(shape) > 2 else 1
        group_size = channels * shape[-2] * shape[-1]
    else:
        group_size = max(wavform.size, 1)
    n_groups = wavform.size // group_size if group_size else 0
    out_dtype = wavform.dtype if wavform.dtype in (np.float32, np.float64) else np.float64
    x = np.ascontiguousarray(wavform, dtype=np.float32).reshape(-1)
    out = np.empty_like(x)
    if x.size:
        eng = get_engine()
        with eng.lock:
            d = eng.buf("ew_in", x.nbytes)
            keep = eng.h2d(d, x)
            L.check(eng.lib.mafe_amplitude_to_db(eng.ctx, d, d, n_groups, group_size, multiplier, amin, offset,
                                                 -1.0 if top_db is None else float(top_db)))
            eng.d2h(out, d)
            eng.sync()
            del keep
    return out.reshape(shape).astype(out_dtype, copy=False)


def dB_to_amplitude(wavform, ref, power):
    """``spectrum.py:93-113``: ``ref * (10 ** (0.1 x)) ** power``."""
    wavform = np.asarray(wavform)
    ref_value = ref(wavform) if callable(ref) else np.abs(ref)
    out_dtype = wavform.dtype if wavform.dtype in (np.float32, np.float64) else np.float64
    x = np.ascontiguousarray(wavform, dtype=np.float32).reshape(-1)
    out = np.empty_like(x)
    if x.size:
        eng = get_engine()
        with eng.lock:
            d = eng.buf("ew_in", x.nbytes)
            keep = eng.h2d(d, x)
            L.check(eng.lib.mafe_db_to_amplitude(eng.ctx, d, d, x.size, float(ref_value), float(power)))
            eng.d2h(out, d)
            eng.sync()
            del keep
    return out.reshape(wavform.shape).astype(out_dtype, copy=False)


# ------------------------------------------------------------------------------------------
def _frames_to_ft(out, lead, frames, dim):
    """[B*T, dim] frame-major -> [..., dim, T] (a transposed VIEW: for one utterance this is the
    F-ordered matrix the reference returns, ``spectrum.py:249-252``)."""
    return np.swapaxes(out.reshape(lead + (frames, dim)), -1, -2)


def stft(
    waveforms,
    n_fft=512,
    win_length=None,
    hop_length=None,
    window="hann",
    center=True,
    pad_mode="constant",
    return_complex=True,
):
    """Short-time Fourier transform, ``spectrum.py:125-278``: complex64 ``[..., 1 + n_fft//2, T]``."""
    waveforms = np.asarray(waveforms)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    fft_window = T.analysis_window(window, win_length, n_fft)
    length = waveforms.shape[-1]
    if center:
        if n_fft > length:
            raise ValueError("n_fft={} is too small for input signal of length={}".format(n_fft, length))
    elif n_fft > length:
        raise ValueError(
            f"n_fft={n_fft} is too large for uncentered analysis of input signal of length={length}")
    if hop_length < 1:
        raise ValueError("Invalid hop_length: {:d}".format(hop_length))
    x, lead = _flatten_batch(waveforms)
    dev_center, dev_pad = center, pad_mode
    if center and pad_mode not in L.PAD:
        # any other np.pad mode (wrap, mean, ...): pad on the host, frame un-centred on the device
        x = np.pad(x, [(0, 0), (n_fft // 2, n_fft // 2)], mode=pad_mode)
        dev_center, dev_pad = False, "constant"
    eng = get_engine()
    plan = eng.plan(n_fft=n_fft, hop=hop_length, center=dev_center, pad_mode=dev_pad if dev_center else "constant",
                    out_kind=L.OUT_COMPLEX, window=fft_window)
    out, fo = eng.run_frontend(plan, x.reshape(-1), _dense_offsets(x.shape[0], x.shape[1]))
    frames = plan.num_frames(x.shape[1])
    n_bins = 1 + n_fft // 2
    stft_matrix = _frames_to_ft(out.view(np.complex64), lead, frames, n_bins)
    if return_complex:
        return stft_matrix
    return np.stack((stft_matrix.real, stft_matrix.imag), -1)


def frame(x, frame_length=2048, hop_length=64):
    """``spectrum.py:281-304`` (host utility kept for the importers in processing.py): ``[.., L]`` ->
    ``[.., frame_length, n_frames]`` float64."""
    if hop_length < 1:
        raise ValueError("Invalid hop_length: {:d}".format(hop_length))
    x = np.asarray(x)
    num_frame = (x.shape[-1] - frame_length) // hop_length + 1
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(max(num_frame, 0))[None, :]
    return x[..., idx].astype(np.float64)


def _pad_shape(y_shift, data_shape):
    """``spectrum.py:307-320``."""
    need_shape = y_shift.shape[-1]
    if need_shape > data_shape:
        return y_shift[..., :data_shape]
    if need_shape < data_shape:
        lengths = [(0, 0)] * y_shift.ndim
        lengths[-1] = (0, data_shape - need_shape)
        return np.pad(y_shift, lengths, mode="constant")
    return y_shift


def _pad_center(data, size, axis=-1):
    """``spectrum.py:323-336``."""
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    if lpad < 0:
        raise ValueError(("Target size ({:d}) must be " "at least input size ({:d})").format(size, n))
    return np.pad(data, lengths)


def istft(
    stft_matrix,
    n_fft=None,
    win_length=None,
    hop_length=None,
    window="hann",
    center=True,
    length=None,
):
    """Inverse STFT, ``spectrum.py:346-474``: float64 ``[..., hop * (T - 1)]`` (or ``length``)."""
    stft_matrix = np.asarray(stft_matrix)
    if n_fft is None:
        n_fft = 2 * (stft_matrix.shape[-2] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    ifft_window = np.ascontiguousarray(T.analysis_window(window, win_length, n_fft), dtype=np.float64)
    if length:
        padded_length = length + int(n_fft) if center else length
        n_frames = min(stft_matrix.shape[-1], int(np.ceil(padded_length / hop_length)))
    else:
        n_frames = stft_matrix.shape[-1]
    lead = stft_matrix.shape[:-2]
    n_bins = stft_matrix.shape[-2]
    if n_bins != n_fft // 2 + 1:
        raise ValueError("stft_matrix has {} rows, expected 1 + n_fft//2 = {}".format(n_bins, n_fft // 2 + 1))
    n_utts = int(np.prod(lead)) if lead else 1
    expected_signal_len = n_fft + hop_length * (n_frames - 1)
    # [.., F, T] -> frame-major [U, T, F] complex64
    z = np.ascontiguousarray(np.swapaxes(stft_matrix[..., :n_frames], -1, -2).reshape((n_utts, n_frames, n_bins)),
                             dtype=np.complex64)
    y64 = np.empty((n_utts, expected_signal_len), dtype=np.float64)
    if z.size:
        eng = get_engine()
        with eng.lock:
            dz = eng.buf("wave", z.nbytes)
            dy = eng.buf("out", y64.nbytes)
            keep = eng.h2d(dz, z)
            L.check(eng.lib.mafe_istft(eng.ctx, dz, n_utts, n_frames, n_fft, hop_length, _vp(ifft_window), dy))
            eng.d2h(y64, dy)
            eng.sync()
            del keep
    y = y64.reshape(lead + (expected_signal_len,))
    if length is None:
        if center:
            y = y[..., int(n_fft // 2): -int(n_fft // 2)]
    else:
        start = int(n_fft // 2) if center else 0
        y = _pad_shape(y[..., start:], data_shape=length)
    return y


def compute_amplitude(waveforms, lengths=None, amp_type="avg", dB=False):
    """``spectrum.py:497-544``.  Host-side scalar utility (one reduction per waveform, not on the
    feature hot path -- SURVEY.md section 8a); kept so processing.py / augment.py importers work."""
    waveforms = np.asarray(waveforms)
    if len(waveforms.shape) == 1:
        waveforms = np.expand_dims(waveforms, 0)
    waveforms = np.abs(waveforms)
    if amp_type == "avg":
        if lengths is None:
            out = waveforms.mean(axis=1, keepdims=True)
        else:
            out = waveforms.sum(axis=1, keepdims=True) / lengths
    elif amp_type == "peak":
        out = waveforms.max(axis=1, keepdims=True)
    else:
        raise TypeError("Unsupported amplitude type {}".format(repr(amp_type)))
    if dB:
        out = 20 * np.log10(out)
        return out.clip(min=-80)
    return out


# ------------------------------------------------------------------------------------------
def _spectrogram_plan(eng, n_fft, win_length, hop_length, window, power, normalized, center, pad_mode, **extra):
    window = WindowType(window)
    pad_mode = BorderType(pad_mode)
    if not float(power) >= 0:
        raise ValueError("power must be non-negative, got {}".format(power))
    w = T.analysis_window(window, win_length, n_fft, coerce=True)
    scale = 1.0 / np.sqrt(np.sum(w ** 2)) if normalized else 1.0
    return eng.plan(n_fft=n_fft, hop=hop_length, center=bool(center), pad_mode=pad_mode.value if center else "constant",
                    window=w, power=power, spec_scale=scale, **extra)


def _run_spec(plan, waveforms, pad, n_fft, center, db_group=L.DBGROUP_NONE, utt_group=None):
    x, lead = _flatten_batch(waveforms)
    if pad > 0:
        x = np.pad(x, [(0, 0), (pad, pad)])
    if x.shape[-1] < n_fft and not center:
        raise ValueError("n_fft={} is too large for input signal of length={}".format(n_fft, x.shape[-1]))
    if center and x.shape[-1] < n_fft // 2 + 1:
        raise ValueError("padding of n_fft // 2 = {} needs a longer input than {}".format(n_fft // 2, x.shape[-1]))
    eng = plan.engine
    out, fo = eng.run_frontend(plan, x.reshape(-1), _dense_offsets(x.shape[0], x.shape[1]), db_group=db_group,
                               utt_group=utt_group)
    return out, lead, plan.num_frames(x.shape[1])


def spectrogram(
    waveforms,
    n_fft=400,
    win_length=None,
    hop_length=None,
    pad=0,
    window="hann",
    power=2.0,
    normalized=False,
    center=True,
    pad_mode="reflect",
    onesided=True,
):
    """``spectrum.py:547-606`` (``msaudio.Spectrogram`` semantics): ``[..., F, T]``."""
    waveforms = np.asarray(waveforms)
    win_length = win_length if win_length else n_fft
    hop_length = hop_length if hop_length else win_length // 2
    window = WindowType(window)
    pad_mode = BorderType(pad_mode)
    eng = get_engine()
    plan = _spectrogram_plan(eng, n_fft, win_length, hop_length, window, power, normalized, center, pad_mode,
                             out_kind=L.OUT_POWER)
    out, lead, frames = _run_spec(plan, waveforms, pad, n_fft, center)
    n_bins = n_fft // 2 + 1
    spec = _frames_to_ft(out, lead, frames, n_bins)
    if not onesided:
        # |X[N-k]| = |X[k]| for real input
        mirror = np.arange(n_fft - n_bins, 0, -1)
        spec = np.concatenate([spec, spec[..., mirror, :]], axis=-2)
    return spec.astype(_out_dtype(waveforms), copy=False)


def _mel_bank(n_fft, n_mels, sample_rate, f_min, f_max, norm, mel_type):
    f_max = f_max if f_max is not None else sample_rate // 2
    if f_min > f_max:
        raise ValueError("f_min ({}) should be no more than f_max ({})".format(f_min, f_max))
    return T.hz_triangle_bank(n_fft // 2 + 1, n_mels, sample_rate, f_min, f_max, norm, mel_type)


def melspectrogram(
    waveforms,
    n_fft=400,
    win_length=None,
    hop_length=None,
    pad=0,
    window="hann",
    power=2.0,
    normalized=False,
    center=True,
    pad_mode="reflect",
    onesided=True,
    n_mels=128,
    sample_rate=16000,
    f_min=0,
    f_max=None,
    norm=NormType.NONE,
    mel_type=MelType.HTK,
):
    """``spectrum.py:609-698``: ``MelScale(Spectrogram(x))`` -> ``[..., n_mels, T]``."""
    waveforms = np.asarray(waveforms)
    win_length = win_length if win_length is not None else n_fft
    hop_length = hop_length if hop_length is not None else win_length // 2
    norm = NormType(norm)
    mel_type = MelType(mel_type)
    window = WindowType(window)
    pad_mode = BorderType(pad_mode)
    if not onesided:
        raise ValueError("MelScale expects n_stft = n_fft // 2 + 1 bins: onesided must be True")
    eng = get_engine()
    bank = _mel_bank(n_fft, n_mels, sample_rate, f_min, f_max, norm, mel_type)
    plan = _spectrogram_plan(eng, n_fft, win_length, hop_length, window, power, normalized, center, pad_mode,
                             out_kind=L.OUT_MEL, mel_fb=bank)
    out, lead, frames = _run_spec(plan, waveforms, pad, n_fft, center)
    return _frames_to_ft(out, lead, frames, n_mels).astype(_out_dtype(waveforms), copy=False)


def magphase(waveform, power, iscomplex=True):
    """``spectrum.py:701-735``.  ``iscomplex=True``: magnitude ``|z| ** power`` and unit phase
    (``0 -> 1+0j``); N-D accepted (the reference handles 2-D only).  ``iscomplex=False``:
    ``msaudio.Magphase`` on ``[..., 2]`` -> (magnitude, angle)."""
    waveform = np.asarray(waveform)
    if power < 0:
        raise ValueError("power must be non-negative, got {}".format(power))
    if iscomplex:
        z = np.ascontiguousarray(waveform, dtype=np.complex64)
        mag_dtype = np.float64 if waveform.dtype == np.complex128 else np.float32
    else:
        if waveform.shape[-1] != 2:
            raise RuntimeError("input tensor is not in shape of <..., complex=2>")
        z = np.ascontiguousarray(waveform, dtype=np.float32).view(np.complex64)[..., 0]
        mag_dtype = _out_dtype(waveform)
    mag = np.empty(z.shape, dtype=np.float32)
    phase = np.empty(z.shape, dtype=np.complex64)
    if z.size:
        eng = get_engine()
        with eng.lock:
            dz = eng.buf("wave", z.nbytes)
            dm = eng.buf("out", mag.nbytes)
            dp = eng.buf("out2", phase.nbytes)
            keep = eng.h2d(dz, z)
            L.check(eng.lib.mafe_magphase(eng.ctx, dz, z.size, float(power), dm, dp))
            eng.d2h(mag, dm)
            eng.d2h(phase, dp)
            eng.sync()
            del keep
    if iscomplex:
        return mag.astype(mag_dtype, copy=False), phase
    return mag.astype(mag_dtype, copy=False), np.arctan2(phase.imag, phase.real).astype(mag_dtype, copy=False)


def melscale(
    spec,
    n_mels=128,
    sample_rate=16000,
    f_min=0,
    f_max=None,
    n_stft=201,
    norm=NormType.NONE,
    mel_type=MelType.HTK,
):
    """``spectrum.py:738-774``: ``[..., n_stft, T]`` -> ``[..., n_mels, T]``."""
    spec = np.asarray(spec)
    f_max = f_max if f_max is not None else sample_rate // 2
    if f_min > f_max:
        raise ValueError("f_min ({}) should be no more than f_max ({})".format(f_min, f_max))
    if spec.ndim < 2 or spec.shape[-2] != n_stft:
        raise RuntimeError("input tensor is not in shape of <..., freq={}, time>, got {}".format(n_stft, spec.shape))
    bank = np.ascontiguousarray(T.hz_triangle_bank(n_stft, n_mels, sample_rate, f_min, f_max, NormType(norm), MelType(mel_type)),
                                dtype=np.float32)
    lead, t = spec.shape[:-2], spec.shape[-1]
    n_mats = int(np.prod(lead)) if lead else 1
    x = np.ascontiguousarray(spec, dtype=np.float32)
    out = np.empty(lead + (n_mels, t), dtype=np.float32)
    if x.size:
        eng = get_engine()
        with eng.lock:
            dx = eng.buf("wave", x.nbytes)
            do = eng.buf("out", out.nbytes)
            keep = eng.h2d(dx, x)
            L.check(eng.lib.mafe_melscale(eng.ctx, dx, do, n_mats, n_stft, t, _vp(bank), n_mels))
            eng.d2h(out, do)
            eng.sync()
            del keep
    return out.astype(_out_dtype(spec), copy=False)


def resynthesize(enhanced_mag, noisy_inputs, normalize_wavs=True):
    """``spectrum.py:777-818``: enhanced magnitude + noisy phase -> waveform (stft/istft on device)."""
    noisy_feats = stft(noisy_inputs, return_complex=False)
    noisy_phase = np.arctan2(noisy_feats[:, :, 1], noisy_feats[:, :, 0])
    pre_stack = np.stack([np.cos(noisy_phase), np.sin(noisy_phase)], axis=-1)
    complex_predictions = np.expand_dims(enhanced_mag, -1) * pre_stack
    result = complex_predictions[:, :, 0] + 1j * complex_predictions[:, :, 1]
    pred_wavs = istft(result)
    if normalize_wavs:
        # processing.normalize(norm="max") (mindaudio/data/processing.py:28-76): peak-normalise
        peak = np.max(np.abs(pred_wavs), axis=-1, keepdims=True)
        pred_wavs = pred_wavs / np.where(peak > 0, peak, 1.0)
    return pred_wavs
