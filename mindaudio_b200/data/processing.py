"""``mindaudio/data/processing.py`` pieces on the feature path (scope rows f2 / f4)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L
from .._engine import get_engine

__all__ = ["sliding_window_cmn", "resample"]


def sliding_window_cmn(x, cmn_window=600, min_cmn_window=100, center=False, norm_vars=False):
    """``processing.py:380-407`` (``msaudio.SlidingWindowCmn``; Kaldi's sliding-window cepstral mean [and variance]
    normalisation as in torchaudio.functional.sliding_window_cmn).  ``x``: ``[..., num_frames, num_feats]``."""
    x = np.asarray(x)
    if x.ndim < 2:
        raise RuntimeError("SlidingWindowCmn: the shape of input tensor does not match the requirement of operator, "
                           "expected at least 2 dimensions, got {}".format(x.ndim))
    if cmn_window < 0 or min_cmn_window < 0:
        raise ValueError("cmn_window and min_cmn_window must be non-negative")
    out_dtype = np.float64 if x.dtype == np.float64 else np.float32
    xf = np.ascontiguousarray(x, dtype=np.float32)
    T, D = xf.shape[-2], xf.shape[-1]
    n_ch = int(np.prod(xf.shape[:-2])) if xf.ndim > 2 else 1
    out = np.empty_like(xf)
    if xf.size:
        eng = get_engine()
        with eng.lock:
            d = eng.buf("wave", xf.nbytes)
            d_out = eng.buf("out", xf.nbytes)
            k = eng.h2d(d, xf)
            L.check(eng.lib.mafe_sliding_window_cmn(eng.ctx, d, d_out, n_ch, T, D, max(int(cmn_window), 1), int(min_cmn_window),
                                                    int(bool(center)), int(bool(norm_vars))))
            eng.d2h(out, d_out)
            eng.sync()
            del k
    return out.astype(out_dtype, copy=False)


def resample(waveform, orig_freq=16000, new_freq=16000, res_type="fft", lowpass_filter_width=6, rolloff=0.99, beta=None):
    """``processing.py:132-186``, the Fourier method (``res_type`` "fft" / "scipy" -> ``scipy.signal.resample`` along
    the last axis to ``ceil(n * new_freq / orig_freq)`` samples), on the device (``mafe_resample_fft``: both DFTs as
    Bluestein transforms in complex128, lengths are arbitrary).  Same dtype out as in, like the reference's
    ``np.asarray(y_hat, dtype=waveform.dtype)``.  ``res_type="minddata"`` (``msaudio.Resample``, a windowed-sinc
    interpolator of the MindSpore wheel) is not on the feature path and not built."""
    if orig_freq == new_freq:
        return waveform
    waveform = np.asarray(waveform)
    ratio = float(new_freq) / orig_freq
    n_in = waveform.shape[-1]
    n_out = int(np.ceil(n_in * ratio))
    if res_type not in ("scipy", "fft"):
        raise NotImplementedError("resample: res_type=%r (msaudio.Resample) is outside the front-end feature path" % (res_type,))
    if np.iscomplexobj(waveform):
        raise NotImplementedError("resample: complex input is outside the front-end feature path")
    if n_out < 1:
        raise ValueError("resample: output length must be positive")
    lead = waveform.shape[:-1]
    rows = int(np.prod(lead)) if lead else 1
    x = np.ascontiguousarray(waveform, dtype=np.float64).reshape(rows, n_in)
    out = np.empty((rows, n_out), dtype=np.float64)
    if rows and n_in:
        eng = get_engine()
        # bound the scratch (two complex128 planes per signal): chunks of rows
        M = 1
        while M < 2 * max(n_in, n_out) - 1:
            M <<= 1
        step = max(1, min(rows, (1 << 25) // M, 65535))     # scratch bound; grid.y limit of the kernels
        with eng.lock:
            for r0 in range(0, rows, step):
                r1 = min(rows, r0 + step)
                need = C.c_size_t()
                L.check(eng.lib.mafe_resample_workspace(r1 - r0, n_in, n_out, C.byref(need)))
                d_x = eng.buf("wave", x[r0:r1].nbytes)
                d_o = eng.buf("out", out[r0:r1].nbytes)
                d_w = eng.buf("work", need.value)
                keep = eng.h2d(d_x, x[r0:r1])
                L.check(eng.lib.mafe_resample_fft(eng.ctx, d_x, r1 - r0, n_in, n_out, d_o, d_w, need.value))
                eng.d2h(out[r0:r1], d_o)
                eng.sync()
                del keep
    elif rows:
        raise ValueError("resample: empty input")
    dt = waveform.dtype if waveform.dtype.kind == "f" else np.float64
    return np.asarray(out.reshape(lead + (n_out,)), dtype=dt)
