"""CMVN family of the reference's pipelines on B200 (numpy in / numpy out):

* ``InputNormalization`` -- examples/ECAPA-TDNN/spec_augment.py:22-70 (sentence-level mean/std)
* ``scalar_norm``        -- examples/deepspeech2/dataset.py:43-47 (log1p + scalar mean/std)
* ``CmvnStats`` / ``compute_cmvn_stats`` / ``save_cmvn_json`` -- examples/conformer/compute_cmvn_stats.py:45-128
* ``load_cmvn`` / ``_load_json_cmvn`` -- mindaudio/utils/load_files.py:9-36
* ``GlobalCMVN``         -- mindaudio/models/layers/cmvn.py:6-36

The per-frame arithmetic (sums, normalisation) runs in libmafe.so; the 2*D+1 statistics are
finalised on the host in float64 exactly as load_files.py does.
"""
from __future__ import annotations

import ctypes as C
import json
import math

import numpy as np

from .. import _lib as L
from .._engine import get_engine

__all__ = ["InputNormalization", "utterance_cmvn", "scalar_norm", "CmvnStats", "compute_cmvn_stats",
           "save_cmvn_json", "load_cmvn", "GlobalCMVN"]


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _run_utt(x, frame_offsets, dim, fn):
    """x: [total_frames, dim] float32 C-contiguous (modified copy returned)."""
    out = np.empty_like(x)
    if x.size:
        eng = get_engine()
        fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        with eng.lock:
            dx = eng.buf("wave", x.nbytes)
            dfo = eng.buf("offsets", fo.nbytes)
            k1, k2 = eng.h2d(dx, x), eng.h2d(dfo, fo)
            fn(eng, dx, dfo, len(fo) - 1)
            eng.d2h(out, dx)
            eng.sync()
            del k1, k2
    return out


def utterance_cmvn(feats, frame_offsets=None, mean_norm=True, std_norm=True):
    """Per-utterance, per-dimension ``(x - mean_t) / std_t`` (population std, NO eps --
    spec_augment.py:43-70).  ``feats``: ``[T, D]`` for one utterance, or the flat ragged
    ``[sum T, D]`` with ``frame_offsets`` (int64 ``[n_utts + 1]``)."""
    feats = np.asarray(feats)
    x = np.ascontiguousarray(feats, dtype=np.float32).reshape((-1, feats.shape[-1]))
    if frame_offsets is None:
        frame_offsets = np.array([0, x.shape[0]], dtype=np.int64)
    dim = x.shape[1]
    out = _run_utt(x, frame_offsets, dim, lambda eng, dx, dfo, n: L.check(
        eng.lib.mafe_cmvn_utt(eng.ctx, dx, dfo, n, dim, int(bool(mean_norm)), int(bool(std_norm)))))
    return out.reshape(feats.shape).astype(feats.dtype if feats.dtype == np.float64 else np.float32, copy=False)


class InputNormalization:
    """``examples/ECAPA-TDNN/spec_augment.py:22-70``: ``construct(x[B, T, D])`` normalises every
    sentence over its (full, padded) time axis when ``norm_type == "sentence"``; other norm types
    leave the input unchanged, as the reference's ``construct`` does."""

    def __init__(self, mean_norm=True, std_norm=True, norm_type="global"):
        self.mean_norm = mean_norm
        self.std_norm = std_norm
        self.norm_type = norm_type
        self.eps = 1e-10  # unused, as in the reference (spec_augment.py:41)

    def construct(self, x_input):
        x_input = np.asarray(x_input)
        if self.norm_type != "sentence":
            return x_input
        b, t = x_input.shape[0], x_input.shape[1]
        flat = x_input.reshape((b * t, -1))
        fo = np.arange(b + 1, dtype=np.int64) * t
        out = utterance_cmvn(flat, fo, self.mean_norm, self.std_norm)
        return out.reshape(x_input.shape)

    __call__ = construct


def scalar_norm(mag, log1p=True):
    """``examples/deepspeech2/dataset.py:43-47``: ``log1p(mag)`` then ``(m - m.mean()) / m.std()``
    over the whole matrix."""
    mag = np.asarray(mag)
    x = np.ascontiguousarray(mag, dtype=np.float32).reshape((-1, 1))
    fo = np.array([0, x.shape[0]], dtype=np.int64)
    out = _run_utt(x, fo, 1, lambda eng, dx, dfo, n: L.check(
        eng.lib.mafe_cmvn_scalar(eng.ctx, dx, dfo, n, 1, int(bool(log1p)))))
    return out.reshape(mag.shape).astype(mag.dtype if mag.dtype == np.float64 else np.float32, copy=False)


class CmvnStats:
    """Running global statistics ``(frame_num, mean_stat[D], var_stat[D])`` in float64, accumulated on
    the device (compute_cmvn_stats.py:61-63, 104-112).  ``allreduce()`` sums them over the ranks
    of a ``torch.distributed`` job (the one collective of the path: 2*D+1 float64 values)."""

    def __init__(self, dim):
        self.dim = dim
        self.mean_stat = np.zeros(dim, dtype=np.float64)
        self.var_stat = np.zeros(dim, dtype=np.float64)
        self.frame_num = 0

    def accumulate(self, feats):
        """feats: ``[T, D]`` (any float dtype) -- one utterance or a flat ragged batch."""
        feats = np.asarray(feats)
        x = np.ascontiguousarray(feats, dtype=np.float32).reshape((-1, self.dim))
        if not x.size:
            return self
        stats = np.zeros(2 * self.dim + 1, dtype=np.float64)
        eng = get_engine()
        with eng.lock:
            dx = eng.buf("wave", x.nbytes)
            ds = eng.buf("stats", stats.nbytes)
            k1, k2 = eng.h2d(dx, x), eng.h2d(ds, stats)
            L.check(eng.lib.mafe_cmvn_stats_accumulate(eng.ctx, dx, x.shape[0], self.dim, ds))
            eng.d2h(stats, ds)
            eng.sync()
            del k1, k2
        self.add_raw(stats)
        return self

    def add_raw(self, stats):
        """stats: float64 ``[2*D+1]`` = (sum x, sum x^2, N) as produced by ``mafe_cmvn_stats_accumulate``."""
        self.mean_stat += stats[: self.dim]
        self.var_stat += stats[self.dim: 2 * self.dim]
        self.frame_num += int(round(float(stats[2 * self.dim])))

    def allreduce(self, group=None):
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self
        packed = np.concatenate([self.mean_stat, self.var_stat, [float(self.frame_num)]])
        dev = torch.device("cuda", get_engine().device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.from_numpy(packed).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        packed = t.cpu().numpy()
        self.mean_stat = packed[: self.dim].copy()
        self.var_stat = packed[self.dim: 2 * self.dim].copy()
        self.frame_num = int(round(float(packed[-1])))
        return self

    def to_dict(self):
        return {"mean_stat": list(self.mean_stat.tolist()), "var_stat": list(self.var_stat.tolist()),
                "frame_num": int(self.frame_num)}

    def mean_istd(self):
        return cmvn_from_stats(self.to_dict())


def compute_cmvn_stats(feats_iter, dim=None):
    """Accumulate over an iterable of ``[T, D]`` feature matrices (compute_cmvn_stats.py:104-112)."""
    stats = None
    for f in feats_iter:
        f = np.asarray(f)
        if stats is None:
            stats = CmvnStats(dim or f.shape[-1])
        stats.accumulate(f)
    return stats


def save_cmvn_json(stats, path):
    """On-disk format of compute_cmvn_stats.py:121-128."""
    d = stats.to_dict() if isinstance(stats, CmvnStats) else stats
    with open(path, "w") as fout:
        fout.write(json.dumps(d))


def cmvn_from_stats(cmvn_stats):
    """``mindaudio/utils/load_files.py:19-28``: float64 ``[2, D]`` = (mean, 1/std), variance floor 1e-20."""
    means = list(cmvn_stats["mean_stat"])
    variance = list(cmvn_stats["var_stat"])
    count = cmvn_stats["frame_num"]
    for i in range(len(means)):
        means[i] /= count
        variance[i] = variance[i] / count - means[i] * means[i]
        if variance[i] < 1.0e-20:
            variance[i] = 1.0e-20
        variance[i] = 1.0 / math.sqrt(variance[i])
    return np.array([means, variance])


def _load_json_cmvn(json_cmvn_file):
    with open(json_cmvn_file) as f:
        return cmvn_from_stats(json.load(f))


def load_cmvn(cmvn_file, is_json):
    """``mindaudio/utils/load_files.py:32-36``."""
    if is_json:
        cmvn = _load_json_cmvn(cmvn_file)
    else:
        raise NotImplementedError("only the json cmvn format exists in the reference")
    return cmvn[0], cmvn[1]


class GlobalCMVN:
    """``mindaudio/models/layers/cmvn.py:6-36``: ``x - mean`` then ``* istd`` (float32), ``x`` is
    ``(batch, max_len, feat_dim)`` or ``(frames, feat_dim)``."""

    def __init__(self, mean, istd, norm_var=True):
        mean, istd = np.asarray(mean), np.asarray(istd)
        assert mean.shape == istd.shape
        self.norm_var = norm_var
        self.mean = np.ascontiguousarray(mean, dtype=np.float32)
        self.istd = np.ascontiguousarray(istd, dtype=np.float32)

    def construct(self, x):
        x = np.asarray(x)
        dim = self.mean.shape[0]
        flat = np.ascontiguousarray(x, dtype=np.float32).reshape((-1, dim))
        out = np.empty_like(flat)
        if flat.size:
            eng = get_engine()
            with eng.lock:
                dx = eng.buf("wave", flat.nbytes)
                dm = eng.buf("cmvn_mean", self.mean.nbytes)
                di = eng.buf("cmvn_istd", self.istd.nbytes)
                k = (eng.h2d(dx, flat), eng.h2d(dm, self.mean), eng.h2d(di, self.istd))
                L.check(eng.lib.mafe_cmvn_apply(eng.ctx, dx, flat.shape[0], dim, dm, di if self.norm_var else None))
                eng.d2h(out, dx)
                eng.sync()
                del k
        return out.reshape(x.shape)

    __call__ = construct
