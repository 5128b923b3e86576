"""SpecAugment-style masking on the device (scope row f4).

* ``spec_aug``          -- the conformer collate's time / frequency masks (``examples/conformer/dataset.py:493-534``).
  The mask positions are drawn on the host with the SAME sequence of ``random`` calls as the reference, so a seeded
  run reproduces the reference bit for bit; the rectangles are zeroed by one kernel over the ragged batch.
* ``frequencymasking`` / ``timemasking`` -- ``mindaudio/data/augment.py:28-98`` (``msaudio.FrequencyMasking`` /
  ``TimeMasking``).  ``iid_masks=False``: the deterministic mask of width ``frequency_mask_param`` at ``mask_start``
  (MindSpore's MaskAlongAxis).  ``iid_masks=True`` draws width and start per example from ``numpy`` (MindSpore's own
  random stream is not reproducible outside MindSpore; parity is in distribution only).
"""
from __future__ import annotations

import ctypes as C
import random as _random

import numpy as np

from .. import _lib as L
from .._engine import get_engine

__all__ = ["spec_aug", "spec_aug_rects", "frequencymasking", "timemasking"]


def spec_aug_rects(shapes, spec_aug_conf, rng=None):
    """Mask rectangles ``[n, 5] = (item, row0, row1, col0, col1)`` for items of shape ``(frames, freq)``, drawn exactly
    like ``dataset.py:515-533`` (time masks first, then frequency masks; three ``randint`` calls per mask)."""
    rng = rng or _random
    num_t_mask = spec_aug_conf.get("num_t_mask", 0)
    num_f_mask = spec_aug_conf.get("num_f_mask", 0)
    max_t = spec_aug_conf.get("max_t", 0)
    max_f = spec_aug_conf.get("max_f", 0)
    rects = []
    for i, (max_frames, max_freq) in enumerate(shapes):
        for _ in range(num_t_mask):
            start = rng.randint(0, max_frames - 1)
            length = rng.randint(1, max_t)
            end = min(max_frames, start + length)
            if rng.randint(1, 100) > 20:
                rects.append((i, start, end, 0, max_freq))
        for _ in range(num_f_mask):
            start = rng.randint(0, max_freq - 1)
            length = rng.randint(1, max_f)
            end = min(max_freq, start + length)
            if rng.randint(1, 100) > 20:
                rects.append((i, 0, max_frames, start, end))
    return np.asarray(rects, dtype=np.int32).reshape(-1, 5)


def _mask_ragged(flat, offsets, dim, rects, value):
    """flat float32 [rows, dim] (modified copy returned), offsets int64 [n + 1], rects int32 [r, 5]."""
    out = np.ascontiguousarray(flat, dtype=np.float32).copy()
    if out.size and len(rects):
        eng = get_engine()
        with eng.lock:
            d = eng.buf("wave", out.nbytes)
            d_fo = eng.buf("stats", offsets.nbytes)
            d_r = eng.buf("aux", rects.nbytes)
            k = (eng.h2d(d, out), eng.h2d(d_fo, offsets), eng.h2d(d_r, rects))
            L.check(eng.lib.mafe_mask_rects(eng.ctx, d, d_fo, len(offsets) - 1, dim, d_r, len(rects), float(value)))
            eng.d2h(out, d)
            eng.sync()
            del k
    return out


def spec_aug(xs, spec_aug_conf, rng=None):
    """``dataset.py:493-534``: list of ``[frames, freq]`` feature matrices -> list of masked matrices (the reference
    masks in place and returns ``xs``; numpy inputs are written back in place here as well when they are float32)."""
    if not xs:
        return xs
    rects = spec_aug_rects([x.shape for x in xs], spec_aug_conf, rng)
    dim = xs[0].shape[1]
    lens = [x.shape[0] for x in xs]
    fo = np.zeros(len(xs) + 1, dtype=np.int64)
    np.cumsum(lens, out=fo[1:])
    flat = np.concatenate([np.asarray(x, dtype=np.float32) for x in xs]) if fo[-1] else np.zeros((0, dim), np.float32)
    out = _mask_ragged(flat, fo, dim, rects, 0.0)
    for i, x in enumerate(xs):
        x[...] = out[fo[i]:fo[i + 1]]
    return xs


def _mask_along(spec, iid_masks, mask_param, mask_start, mask_value, axis, rng):
    spec = np.asarray(spec)
    if spec.ndim < 2:
        raise RuntimeError("input tensor is not in shape of <..., freq, time>")
    size = spec.shape[axis]
    if mask_param < 0 or mask_param > size:
        raise ValueError("mask_param should be in [0, {}], got {}".format(size, mask_param))
    if mask_start < 0 or mask_start > size - mask_param:
        raise ValueError("mask_start should be in [0, {}], got {}".format(size - mask_param, mask_start))
    out_dtype = np.float64 if spec.dtype == np.float64 else np.float32
    F, T = spec.shape[-2], spec.shape[-1]
    n = int(np.prod(spec.shape[:-2])) if spec.ndim > 2 else 1
    rects = []
    if iid_masks:
        rng = rng or np.random.default_rng()
        for i in range(n):
            w = int(rng.integers(0, mask_param + 1)) if mask_param > 0 else 0
            s = int(rng.integers(0, size - w + 1))
            rects.append((i, s, s + w, 0, T) if axis == -2 else (i, 0, F, s, s + w))
    elif mask_param > 0:
        for i in range(n):
            rects.append((i, mask_start, mask_start + mask_param, 0, T) if axis == -2 else
                         (i, 0, F, mask_start, mask_start + mask_param))
    rects = np.asarray(rects, dtype=np.int32).reshape(-1, 5)
    fo = np.arange(n + 1, dtype=np.int64) * F
    out = _mask_ragged(spec.reshape(-1, T), fo, T, rects, mask_value)
    return out.reshape(spec.shape).astype(out_dtype, copy=False)


def frequencymasking(waveform, iid_masks=False, frequency_mask_param=0, mask_start=0, mask_value=0.0, rng=None):
    """``augment.py:28-62``: mask along the frequency axis (dim -2) of ``[..., freq, time]``."""
    return _mask_along(waveform, iid_masks, frequency_mask_param, mask_start, mask_value, -2, rng)


def timemasking(waveform, iid_masks=False, frequency_mask_param=0, mask_start=0, mask_value=0.0, rng=None):
    """``augment.py:65-98`` (the parameter really is called ``frequency_mask_param`` there): mask along time (dim -1)."""
    return _mask_along(waveform, iid_masks, frequency_mask_param, mask_start, mask_value, -1, rng)
