"""``mindaudio/data/augment.py`` pieces built on the STFT path (scope row f2: callers of ``istft``) plus the
spectrogram maskings (row f4).

``time_stretch`` (augment.py:795-825) runs STFT -> phase vocoder -> ISTFT without leaving the GPU: one H2D of the
waveform, one D2H of the stretched signal.  ``_phase_vocoder`` (:828-871) keeps the reference's arithmetic: the phase
accumulator is a float32 array updated in float64 (numpy's in-place ``+=``), magnitudes and output are complex64.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L
from .. import _tables as T
from .._engine import get_engine
from .masking import frequencymasking, timemasking  # noqa: F401
from .processing import resample
from .spectrum import _dense_offsets, _flatten_batch, _pad_shape

__all__ = ["time_stretch", "pitch_shift", "speed_perturb", "_phase_vocoder", "frequencymasking", "timemasking"]


def _vocoder_tables(n_frames, n_bins, rate, hop_length):
    n_steps = len(np.arange(0, n_frames, rate, dtype=np.float64))
    phi = np.ascontiguousarray(np.linspace(0, np.pi * hop_length, n_bins), dtype=np.float64)
    return n_steps, phi


def _phase_vocoder(matrix, rate, hop_length=None, n_fft=None):
    """``augment.py:828-871``: complex ``[..., F, T]`` -> ``[..., F, ceil(T / rate)]`` (same dtype)."""
    matrix = np.asarray(matrix)
    if n_fft is None:
        n_fft = 2 * (matrix.shape[-2] - 1)
    if hop_length is None:
        hop_length = int(n_fft // 4)
    F, Tn = matrix.shape[-2], matrix.shape[-1]
    lead = matrix.shape[:-2]
    n = int(np.prod(lead)) if lead else 1
    n_steps, phi = _vocoder_tables(Tn, F, rate, hop_length)
    z = np.ascontiguousarray(np.swapaxes(matrix, -1, -2).reshape((n, Tn, F)), dtype=np.complex64)
    out = np.empty((n, n_steps, F), dtype=np.complex64)
    if out.size:
        eng = get_engine()
        with eng.lock:
            dz = eng.buf("wave", max(z.nbytes, 16))
            dp = eng.buf("stats", phi.nbytes)
            do = eng.buf("out", out.nbytes)
            keep = (eng.h2d(dz, z), eng.h2d(dp, phi))
            L.check(eng.lib.mafe_phase_vocoder(eng.ctx, dz, n, Tn, F, float(rate), dp, n_steps, do))
            eng.d2h(out, do)
            eng.sync()
            del keep
    out = np.swapaxes(out.reshape(lead + (n_steps, F)), -1, -2)
    return out.astype(matrix.dtype, copy=False) if np.iscomplexobj(matrix) else out


def time_stretch(waveforms, rate=None):
    """``augment.py:795-825``: ``stft`` (512 / 128, hann, centred) -> phase vocoder -> ``istft`` to
    ``round(len / rate)`` samples; float64 ``[..., round(len / rate)]`` like the reference's ``istft``."""
    if rate <= 0:
        raise ValueError("rate must be a positive number")
    waveforms = np.asarray(waveforms)
    n_fft, hop = 512, 128
    x, lead = _flatten_batch(waveforms)
    if n_fft > x.shape[-1]:   # the reference goes through stft(center=True), which refuses these (spectrum.py:182-187)
        raise ValueError("n_fft={} is too small for input signal of length={}".format(n_fft, x.shape[-1]))
    length_stretch = int(round(waveforms.shape[-1] / rate))
    eng = get_engine()
    win = T.analysis_window("hann", n_fft, n_fft)
    plan = eng.plan(n_fft=n_fft, hop=hop, center=True, pad_mode="constant", out_kind=L.OUT_COMPLEX, window=win)
    n_utts, n_bins = x.shape[0], n_fft // 2 + 1
    n_frames = plan.num_frames(x.shape[1])
    n_steps, phi = _vocoder_tables(n_frames, n_bins, rate, hop)
    # istft(spec_stretch, length=length_stretch): only the frames that reach the requested length are used
    n_used = min(n_steps, int(np.ceil((length_stretch + n_fft) / hop)))
    sig_len = n_fft + hop * (n_used - 1)
    y64 = np.empty((n_utts, max(sig_len, 0)), dtype=np.float64)
    if n_used > 0 and n_utts:
        xf = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        with eng.lock:
            b = eng.batch(plan, _dense_offsets(n_utts, x.shape[1]))
            try:
                dw = eng.buf("wave", xf.nbytes)
                ds = eng.buf("out", n_utts * n_frames * n_bins * 8)
                dv = eng.buf("pad", n_utts * n_used * n_bins * 8)
                dp = eng.buf("stats", phi.nbytes)
                dy = eng.buf("aux", y64.nbytes)
                keep = (eng.h2d(dw, xf), eng.h2d(dp, phi))
                L.check(eng.lib.mafe_frontend_run(eng.ctx, plan.h, b.h, dw, L.WAVE_F32, 1.0, ds, L.DBGROUP_NONE))
                L.check(eng.lib.mafe_phase_vocoder(eng.ctx, ds, n_utts, n_frames, n_bins, float(rate), dp, n_used, dv))
                w64 = np.ascontiguousarray(win, dtype=np.float64)
                L.check(eng.lib.mafe_istft(eng.ctx, dv, n_utts, n_used, n_fft, hop, w64.ctypes.data_as(C.c_void_p), dy))
                eng.d2h(y64, dy)
                eng.sync()
                del keep
            finally:
                b.close()
    y = y64.reshape(lead + (y64.shape[-1],))
    return _pad_shape(y[..., n_fft // 2:], data_shape=length_stretch)


def pitch_shift(waveforms, sr, n_steps, bins_per_octave=12):
    """``augment.py:874-901``: ``time_stretch`` by ``2 ** (-n_steps / bins_per_octave)``, then Fourier resampling from
    ``sr / rate`` back to ``sr`` (``processing.resample``), then crop / zero-pad to the length of the STRETCHED signal
    (the reference passes ``waveforms_stretch.shape[-1]`` to ``_pad_shape``, not the input length)."""
    rate = 2.0 ** (-float(n_steps) / bins_per_octave)
    stretched = time_stretch(waveforms, rate=rate)
    y_shift = resample(stretched, orig_freq=float(sr) / rate, new_freq=sr)
    return _pad_shape(y_shift, data_shape=stretched.shape[-1])


def speed_perturb(waveform, orig_freq, speeds=(90, 100, 110), perturb_prob=1.0):
    """``augment.py:601-638``: with probability ``perturb_prob`` resample the batch from ``orig_freq`` to
    ``orig_freq * speed // 100`` for a ``speed`` drawn from ``speeds`` (same ``np.random`` call sequence as the
    reference: ``rand(1)`` then ``randint(0, len(speeds), (1,))``); the Fourier ``resample`` runs on the device."""
    if np.random.rand(1) > perturb_prob:
        return np.asarray(waveform).copy()
    samp_index = np.random.randint(0, len(speeds), (1,))[0]
    speed = speeds[samp_index]
    new_freq = orig_freq * speed // 100
    return resample(waveform, orig_freq, new_freq)
