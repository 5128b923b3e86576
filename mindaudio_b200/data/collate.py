"""Collate step after the front-end ("next" row f3 of the scope table): ragged feature matrices -> the padded batch
the model consumes.

Mirrors ``mindaudio/utils/common.py:10-52`` (``pad_sequence``) and ``mindaudio/utils/mask.py`` (``make_pad_mask``) as
used by the conformer collate (``examples/conformer/dataset.py:563-569, 616-621``).  Float feature sequences are
padded on the GPU (``mafe_pad_sequence``); label sequences (1-D integer arrays of a few dozen entries) are host work
in the reference and stay host work here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L
from .._engine import get_engine

__all__ = ["pad_sequence", "make_pad_mask", "pad_features"]

IGNORE_ID = -1


def make_pad_mask(lengths, max_len=0):
    """``mindaudio/utils/mask.py``: bool ``[B, max_len]``, True at PADDED positions."""
    lengths = np.asarray(lengths)
    max_len = max_len if max_len > 0 else int(lengths.max())
    return np.arange(max_len)[None, :] >= lengths[:, None]


def pad_features(flat, frame_offsets, max_len=None, padding_value=0.0, batch_first=True, with_mask=False):
    """Ragged ``[sum T_i, D]`` float32 features + int64 offsets ``[B + 1]`` -> ``[B, max_len, D]`` float32
    (``[max_len, B, D]`` if not ``batch_first``); optionally also ``xs_masks [B, 1, max_len]`` float32
    (1 = frame, 0 = padding; dataset.py:620-621)."""
    flat = np.ascontiguousarray(flat, dtype=np.float32)
    fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    n, dim = len(fo) - 1, (flat.shape[1] if flat.ndim == 2 else 1)
    lens = np.diff(fo)
    if max_len is None:
        max_len = int(lens.max()) if n else 0
    shape = (n, max_len, dim) if batch_first else (max_len, n, dim)
    out = np.empty(shape, dtype=np.float32)
    mask = np.empty((n, 1, max_len), dtype=np.float32) if with_mask else None
    if out.size:
        eng = get_engine()
        with eng.lock:
            d_in = eng.buf("wave", max(flat.nbytes, 16))
            d_fo = eng.buf("stats", fo.nbytes)
            d_out = eng.buf("out", out.nbytes)
            d_mask = eng.buf("aux", mask.nbytes) if with_mask else None
            k1, k2 = eng.h2d(d_in, flat), eng.h2d(d_fo, fo)
            L.check(eng.lib.mafe_pad_sequence(eng.ctx, d_in, d_fo, n, dim, max_len, float(padding_value), int(bool(batch_first)),
                                              d_out, d_mask))
            eng.d2h(out, d_out)
            if with_mask:
                eng.d2h(mask, d_mask)
            eng.sync()
            del k1, k2
    return (out, mask) if with_mask else out


def pad_sequence(sequences, batch_first=True, padding_value=0, padding_max_len=None, atype=np.int32):
    """``mindaudio/utils/common.py:10-52``: same signature, defaults and result (dtype ``atype``).

    Sequences are truncated to ``padding_max_len``.  2-D float32 sequences of equal trailing size go through the GPU
    kernel; anything else (integer label sequences, 1-D arrays, other dtypes) is assembled on the host exactly like
    the reference does."""
    trailing = tuple(sequences[0].shape[1:])
    max_len = padding_max_len if padding_max_len is not None else max(s.shape[0] for s in sequences)
    dev_ok = (np.dtype(atype) == np.float32 and len(trailing) == 1 and
              all(s.ndim == 2 and s.shape[1:] == trailing for s in sequences))
    if dev_ok:
        lens = [min(s.shape[0], max_len) for s in sequences]
        fo = np.zeros(len(sequences) + 1, dtype=np.int64)
        np.cumsum(lens, out=fo[1:])
        flat = np.concatenate([np.asarray(s[:n], dtype=np.float32) for s, n in zip(sequences, lens)]) if fo[-1] else \
            np.zeros((0, trailing[0]), np.float32)
        return pad_features(flat, fo, max_len, float(padding_value), batch_first)
    dims = ((len(sequences), max_len) if batch_first else (max_len, len(sequences))) + trailing
    out = np.full(dims, fill_value=padding_value).astype(atype)
    for i, seq in enumerate(sequences):
        n = min(seq.shape[0], max_len)
        if batch_first:
            out[i, :n, ...] = seq[:n]
        else:
            out[:n, i, ...] = seq[:n]
    return out
