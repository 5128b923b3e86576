"""CPU arm of bench.py -- TEST/BENCH INFRASTRUCTURE (see oracle/__init__.py).

``conformer_frontend_refstyle`` follows the reference's conformer front-end STEP BY STEP the way
the reference executes it (examples/conformer/dataset.py:117-168): whole-signal pre-emphasis, a
python loop over frames that copies and windows each frame, one scalar mean over the frame
matrix, ``np.fft.rfft(n=512)``, a dense float64 ``[T,257] @ [257,80]`` product, log; followed by
the per-utterance mean/std normalisation of examples/ECAPA-TDNN/spec_augment.py:43-70.  It is
the honest "what the reference costs on this host" number (kind = "port": the reference itself
cannot travel to the GPU box).  ``run_pool`` fans utterances out over a ``multiprocessing.Pool``
the way the reference does (``mp.Pool(8)``, dataset.py:449,479).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import restated as R

_BANK = None
_WIN = None


def _tables():
    global _BANK, _WIN
    if _BANK is None:
        _BANK = R.kaldi_mel_banks(80, 512, 16000.0, 20.0, 8000.0)
        _WIN = R.povey_window(400)
    return _BANK, _WIN


def conformer_frontend_refstyle(wav, utt_cmvn=True):
    bank, win = _tables()
    x = np.asarray(wav, dtype=np.float64)
    sig = np.append(x[0], x[1:] - 0.97 * x[:-1])                 # dataset.py:117-119
    n_frames = int(np.floor((sig.size - 400) / 160) + 1)          # dataset.py:127
    frames = np.zeros((n_frames, 400))
    for i in range(n_frames):                                     # dataset.py:129-131 (per-frame python loop)
        frames[i, :] = sig[i * 160: i * 160 + 400]
        frames[i, :] = frames[i, :] * win
    frames -= np.mean(frames)                                     # dataset.py:165
    spec = np.abs(np.fft.rfft(frames, n=512)) ** 2                # dataset.py:137-138
    feats = np.dot(spec, bank.T)                                  # dataset.py:153
    feats = np.where(feats == 0, np.finfo(float).eps, feats)
    feats = np.log(feats)
    if utt_cmvn:                                                  # spec_augment.py:43-70
        feats = (feats - np.mean(feats, axis=0)) / np.std(feats, axis=0)
    return feats.astype(np.float32)                               # pad_sequence casts to float32 (common.py:10-52)


def _work(args):
    seed, n = args
    rng = np.random.default_rng(seed)
    wav = np.round(np.clip(0.05 * rng.standard_normal(n), -1.0, 1.0) * 32768.0)
    t0 = time.perf_counter()
    out = conformer_frontend_refstyle(wav)
    return time.perf_counter() - t0, out.shape[0]


def run_pool(lengths, seed=3, procs=None):
    """Times the front-end over utterances of the given lengths on ``procs`` host processes.
    Waveform synthesis is excluded from the timed region per worker; the wall clock of the
    ``Pool.map`` (which includes it) is returned too.  -> dict(audio_s, wall_s, cpu_s, frames, procs)"""
    procs = procs or os.cpu_count() or 1
    jobs = [(seed * 1000003 + i, int(n)) for i, n in enumerate(lengths)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_work, jobs[: procs])                # warm the workers (imports, tables)
        t0 = time.perf_counter()
        res = pool.map(_work, jobs, chunksize=max(1, len(jobs) // (procs * 8)))
        wall = time.perf_counter() - t0
    cpu_s = float(sum(r[0] for r in res))
    return {"audio_s": float(sum(lengths)) / 16000.0, "wall_s": wall, "cpu_s": cpu_s,
            "frames": int(sum(r[1] for r in res)), "procs": procs,
            # throughput of the compute alone if the workers were perfectly parallel
            "parallel_compute_s": cpu_s / procs}
