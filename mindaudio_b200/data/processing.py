"""``mindaudio/data/processing.py`` pieces on the feature path (scope row f4)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L
from .._engine import get_engine

__all__ = ["sliding_window_cmn"]


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
