"""CPU arm of bench.py -- TEST/BENCH INFRASTRUCTURE (see oracle/__init__.py).

``conformer_frontend_refstyle`` follows the reference's conformer front-end STEP BY STEP the way
the reference executes it (examples/conformer/dataset.py:117-168): whole-signal pre-emphasis, a
python loop over frames that copies and windows each frame, one scalar mean over the frame
matrix, ``np.fft.rfft(n=512)``, a dense float64 ``[T,257] @ [257,80]`` product, log; followed by
the per-utterance mean/std normalisation of examples/ECAPA-TDNN/spec_augment.py:43-70.  It is
the honest "what the reference costs on this host" number (kind = "port": the reference itself
cannot travel to the GPU box).  ``run_pool`` fans utterances out over a ``multiprocessing.Pool``
the way the reference does (``mp.Pool(8)``, dataset.py:449,479).

Threading: every pool worker runs its BLAS / OpenMP single-threaded (``threadpoolctl`` in the worker
initialiser plus the ``*_NUM_THREADS`` variables): ``procs`` workers whose ``np.dot`` each spawn
``cpu_count`` OpenBLAS threads oversubscribe the host ~``procs``-fold (round-1 finding: 22x too slow
whenever the launcher had not exported ``OMP_NUM_THREADS=1`` already).  ``run_single`` is the
one-process figure (BLAS threads as numpy finds them) BASELINE.md asks for beside the pool figure.
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


_THREAD_VARS = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS")
_LIMITER = None


def _worker_init():
    """One BLAS / OpenMP thread per pool worker (the pool is the parallelism)."""
    global _LIMITER
    for k in _THREAD_VARS:
        os.environ[k] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _LIMITER = threadpool_limits(limits=1)      # kept alive for the life of the worker
    except Exception:                               # threadpoolctl missing: the variables above still cover
        _LIMITER = None                             # libraries loaded after the fork
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass


def blas_threads():
    """BLAS threads of THIS process as threadpoolctl sees them (for the report)."""
    try:
        from threadpoolctl import threadpool_info
        return max([int(i.get("num_threads", 1)) for i in threadpool_info()] or [1])
    except Exception:
        return None


def run_single(lengths, seed=3):
    """The same work in ONE process (no pool), BLAS threads left as numpy configures them."""
    t_cpu = 0.0
    frames = 0
    t0 = time.perf_counter()
    for i, n in enumerate(lengths):
        dt, f = _work((seed * 1000003 + i, int(n)))
        t_cpu += dt
        frames += f
    return {"audio_s": float(sum(lengths)) / 16000.0, "wall_s": time.perf_counter() - t0, "cpu_s": t_cpu,
            "frames": frames, "procs": 1, "blas_threads": blas_threads()}


def run_pool(lengths, seed=3, procs=None):
    """Times the front-end over utterances of the given lengths on ``procs`` host processes.
    Waveform synthesis is excluded from the timed region per worker; the wall clock of the
    ``Pool.map`` (which includes it) is returned too.  -> dict(audio_s, wall_s, cpu_s, frames, procs)"""
    procs = procs or os.cpu_count() or 1
    jobs = [(seed * 1000003 + i, int(n)) for i, n in enumerate(lengths)]
    ctx = mp.get_context("fork")
    saved = {k: os.environ.get(k) for k in _THREAD_VARS}
    for k in _THREAD_VARS:                            # inherited by the forked workers; restored below
        os.environ[k] = "1"
    try:
        pool = ctx.Pool(procs, initializer=_worker_init)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    with pool:
        pool.map(_work, jobs[: procs])                # warm the workers (imports, tables)
        t0 = time.perf_counter()
        res = pool.map(_work, jobs, chunksize=max(1, len(jobs) // (procs * 8)))
        wall = time.perf_counter() - t0
    cpu_s = float(sum(r[0] for r in res))
    return {"audio_s": float(sum(lengths)) / 16000.0, "wall_s": wall, "cpu_s": cpu_s,
            "frames": int(sum(r[1] for r in res)), "procs": procs, "blas_threads_per_proc": 1,
            # throughput of the compute alone if the workers were perfectly parallel
            "parallel_compute_s": cpu_s / procs}
