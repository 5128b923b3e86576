"""Host runtime above the C ABI: lazy per-process context, plan cache, device buffers, and the
numpy-in / numpy-out execution of one front-end call.  Pure ctypes + numpy (no torch).

Fork safety (callers run in ``mp.Pool`` workers): nothing here touches
CUDA until the first op of a process; a forked child builds its own context on first use.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _lib as L


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class Plan:
    """Immutable front-end configuration living on the device (mafe_plan)."""

    def __init__(self, engine, handle, desc_kw):
        self.engine, self.h, self.kw = engine, handle, desc_kw
        lib = engine.lib
        self.out_dim = lib.mafe_plan_out_dim(handle)
        self.is_fast = bool(lib.mafe_plan_is_fast(handle))

    def num_frames(self, n):
        return int(self.engine.lib.mafe_plan_num_frames(self.h, int(n)))


class Engine:
    def __init__(self, device=None):
        self.lib = L.load()
        if device is None:
            device = int(os.environ.get("MAFE_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.device = device
        h = C.c_void_p()
        L.check(self.lib.mafe_ctx_create(device, C.byref(h)))
        self.ctx = h
        self.pid = os.getpid()
        self.lock = threading.RLock()
        self._plans = {}
        self._batch_pool = []     # handles of closed batches, re-laid by mafe_batch_refill (Engine.batch)
        self._bufs = {}
        self._pinned = {}

    # ---- plumbing ----
    def sync(self):
        L.check(self.lib.mafe_ctx_sync(self.ctx))

    def set_stream(self, cuda_stream):
        L.check(self.lib.mafe_ctx_set_stream(self.ctx, C.c_void_p(cuda_stream) if cuda_stream else None))

    @property
    def launch_count(self):
        return int(self.lib.mafe_ctx_launch_count(self.ctx))

    @property
    def sm_count(self):
        return int(self.lib.mafe_ctx_sm_count(self.ctx))

    def profile(self, enable):
        L.check(self.lib.mafe_ctx_profile_enable(self.ctx, int(bool(enable))))

    def profile_reset(self):
        L.check(self.lib.mafe_ctx_profile_reset(self.ctx))

    def profile_read(self, which):
        """(summed device milliseconds, launches) of kernel class ``which`` (L.PROF_*)."""
        ms, n = C.c_double(), C.c_int64()
        L.check(self.lib.mafe_ctx_profile_read(self.ctx, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def fp32_fma_peak(self):
        """Measured FP32 FMA peak of this GPU in TFLOP/s (register-resident FFMA microbenchmark)."""
        v = C.c_double()
        L.check(self.lib.mafe_fp32_fma_peak(self.ctx, C.byref(v)))
        return v.value

    def buf(self, name, nbytes):
        """Grow-only named device buffer (numpy API calls are synchronous, so reuse is safe)."""
        cur = self._bufs.get(name)
        if cur is not None and cur[1] >= nbytes:
            return cur[0]
        if cur is not None:
            L.check(self.lib.mafe_device_free(self.ctx, cur[0]))
            del self._bufs[name]
        p = C.c_void_p()
        cap = max(int(nbytes * 1.25), 1 << 16)
        L.check(self.lib.mafe_device_malloc(self.ctx, cap, C.byref(p)))
        self._bufs[name] = (p, cap)
        return p

    def pinned(self, name, nbytes):
        """Grow-only named pinned host buffer (``(address, capacity)``): staging for asynchronous copies."""
        cur = self._pinned.get(name)
        if cur is not None and cur[1] >= nbytes:
            return cur
        if cur is not None:
            L.check(self.lib.mafe_pinned_free(self.ctx, cur[0]))
            del self._pinned[name]
        p = C.c_void_p()
        cap = max(int(nbytes * 1.25), 1 << 16)
        L.check(self.lib.mafe_pinned_malloc(self.ctx, cap, C.byref(p)))
        self._pinned[name] = (p, cap)
        return self._pinned[name]

    def h2d_raw(self, dev, host_ptr, nbytes):
        """Asynchronous copy from a raw host address (pinned memory keeps it asynchronous)."""
        L.check(self.lib.mafe_memcpy_h2d(self.ctx, dev, host_ptr, int(nbytes)))

    def h2d(self, dev, arr):
        arr = np.ascontiguousarray(arr)
        L.check(self.lib.mafe_memcpy_h2d(self.ctx, dev, arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        return arr  # keep alive until sync

    def d2h(self, arr, dev):
        assert arr.flags["C_CONTIGUOUS"]
        L.check(self.lib.mafe_memcpy_d2h(self.ctx, arr.ctypes.data_as(C.c_void_p), dev, arr.nbytes))

    def d2h_staged(self, arr, dev):
        """Large results: device -> pinned staging at full PCIe rate, then a multi-threaded copy into ``arr`` (synchronous)."""
        assert arr.flags["C_CONTIGUOUS"]
        L.check(self.lib.mafe_memcpy_d2h_staged(self.ctx, arr.ctypes.data_as(C.c_void_p), dev, arr.nbytes))

    def h2d_gather(self, dev, arrays):
        """Upload a list of C-contiguous arrays back to back at ``dev`` (the library's host threads gather them into pinned
        staging, one async copy uploads it); the arrays may be released when this returns."""
        n = len(arrays)
        if n == 0:
            return
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
        sizes = np.fromiter((a.nbytes for a in arrays), dtype=np.int64, count=n)
        L.check(self.lib.mafe_memcpy_h2d_gather(self.ctx, dev, ptrs, sizes.ctypes.data_as(C.c_void_p), n))

    # ---- plans ----
    def plan(self, *, n_fft, frame_len=None, hop, center, pad_mode="constant", out_kind, window, preemph=0.0,
             remove_frame_mean=False, dither=0.0, dither_seed=0, power=2.0, spec_scale=1.0, mel_fb=None,
             log_kind=L.LOG_NONE, log_arg=0.0, log_mult=10.0, log_offset=0.0, top_db=-1.0, dct=None,
             utt_cmvn_mean=False, utt_cmvn_std=False, allow_fast_path=True, utt_scalar_norm=False):
        frame_len = n_fft if frame_len is None else frame_len
        window = np.ascontiguousarray(window, dtype=np.float32)
        mel_fb = None if mel_fb is None else np.ascontiguousarray(mel_fb, dtype=np.float32)
        dct = None if dct is None else np.ascontiguousarray(dct, dtype=np.float32)
        key = (n_fft, frame_len, hop, bool(center), pad_mode, out_kind, window.tobytes(), float(preemph),
               bool(remove_frame_mean), float(dither), int(dither_seed), float(power), float(spec_scale),
               None if mel_fb is None else (mel_fb.shape, mel_fb.tobytes()), log_kind, float(log_arg),
               float(log_mult), float(log_offset), float(top_db), None if dct is None else (dct.shape, dct.tobytes()),
               bool(allow_fast_path), bool(utt_cmvn_mean), bool(utt_cmvn_std), bool(utt_scalar_norm))
        with self.lock:
            p = self._plans.get(key)
            if p is not None:
                return p
            if window.shape != (frame_len,):
                raise ValueError("window must have frame_len=%d entries, got %s" % (frame_len, window.shape))
            d = L.FrontendDesc()
            d.n_fft, d.frame_len, d.hop, d.center = n_fft, frame_len, hop, int(bool(center))
            if pad_mode not in L.PAD:
                raise ValueError("unsupported pad_mode %r (supported: %s)" % (pad_mode, sorted(L.PAD)))
            d.pad_mode, d.out_kind, d.window = L.PAD[pad_mode], out_kind, _fptr(window)
            d.preemph, d.remove_frame_mean = float(preemph), int(bool(remove_frame_mean))
            d.dither, d.dither_seed = float(dither), int(dither_seed) & 0xFFFFFFFFFFFFFFFF
            d.power, d.spec_scale = float(power), float(spec_scale)
            if mel_fb is not None:
                if mel_fb.ndim != 2 or mel_fb.shape[1] != n_fft // 2 + 1:
                    raise ValueError("mel_fb must be [n_mels, n_fft//2+1], got %s" % (mel_fb.shape,))
                d.n_mels, d.mel_fb = mel_fb.shape[0], _fptr(mel_fb)
            d.log_kind, d.log_arg, d.log_mult = log_kind, float(log_arg), float(log_mult)
            d.log_offset, d.top_db = float(log_offset), float(top_db)
            if dct is not None:
                d.n_mfcc, d.dct = dct.shape[1], _fptr(dct)
            d.utt_cmvn_mean, d.utt_cmvn_std = int(bool(utt_cmvn_mean)), int(bool(utt_cmvn_std))
            d.allow_fast_path = int(bool(allow_fast_path))
            d.utt_scalar_norm = int(bool(utt_scalar_norm))
            h = C.c_void_p()
            L.check(self.lib.mafe_plan_create(self.ctx, C.byref(d), C.byref(h)))
            p = Plan(self, h, dict(n_fft=n_fft, frame_len=frame_len, hop=hop, center=center, out_kind=out_kind))
            self._plans[key] = p
            return p

    # ---- ragged batch ----
    def batch(self, plan, sample_offsets, utt_group=None):
        """Batch layout for ``sample_offsets``.  Closed batches go back to a small pool and are re-laid with
        ``mafe_batch_refill`` (their device tables only grow): no cudaMalloc / cudaFree per call."""
        so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
        ug = None if utt_group is None else np.ascontiguousarray(utt_group, dtype=np.int32)
        ugp = None if ug is None else ug.ctypes.data_as(C.c_void_p)
        with self.lock:
            h = self._batch_pool.pop() if self._batch_pool else None
        if h is not None:
            try:
                L.check(self.lib.mafe_batch_refill(self.ctx, plan.h, h, so.ctypes.data_as(C.c_void_p), len(so) - 1, ugp))
            except Exception:
                self.lib.mafe_batch_destroy(h)
                raise
            return Batch(self, h, len(so) - 1)
        h = C.c_void_p()
        L.check(self.lib.mafe_batch_create(self.ctx, plan.h, so.ctypes.data_as(C.c_void_p), len(so) - 1, ugp, C.byref(h)))
        return Batch(self, h, len(so) - 1)

    def _recycle_batch(self, h):
        with self.lock:
            if len(self._batch_pool) < 4:
                self._batch_pool.append(h)
                return
        self.lib.mafe_batch_destroy(h)

    # ---- numpy in / numpy out ----
    def run_frontend(self, plan, flat_wave, sample_offsets, wave_scale=1.0, db_group=L.DBGROUP_NONE, utt_group=None):
        """flat_wave: 1-D float32 / int16 array; returns (out [total_frames, out_dim] float32, frame_offsets)."""
        with self.lock:
            if flat_wave.dtype == np.int16:
                wdt = L.WAVE_I16
            else:
                flat_wave = np.ascontiguousarray(flat_wave, dtype=np.float32)
                wdt = L.WAVE_F32
            b = self.batch(plan, sample_offsets, utt_group)
            try:
                out = np.empty((b.total_frames, plan.out_dim), dtype=np.float32)
                if b.total_frames:
                    dw = self.buf("wave", flat_wave.nbytes)
                    do = self.buf("out", out.nbytes)
                    big = 4 << 20     # large arrays go through the pinned staging buffers (multi-threaded host copies)
                    keep = None
                    if flat_wave.nbytes >= big:
                        self.h2d_gather(dw, [flat_wave])
                    else:
                        keep = self.h2d(dw, flat_wave)
                    L.check(self.lib.mafe_frontend_run(self.ctx, plan.h, b.h, dw, wdt, float(wave_scale), do, db_group))
                    if out.nbytes >= big:
                        self.d2h_staged(out, do)
                    else:
                        self.d2h(out, do)
                    self.sync()
                    del keep
                return out, b.frame_offsets
            finally:
                b.close()


class Batch:
    def __init__(self, engine, handle, n_utts):
        self.engine, self.h, self.n_utts = engine, handle, n_utts
        lib = engine.lib
        self.total_frames = int(lib.mafe_batch_total_frames(handle))
        self.total_samples = int(lib.mafe_batch_total_samples(handle))
        fo = np.empty(n_utts + 1, dtype=np.int64)
        L.check(lib.mafe_batch_frame_offsets(handle, fo.ctypes.data_as(C.c_void_p)))
        self.frame_offsets = fo
        self.frame_offsets_dev = lib.mafe_batch_frame_offsets_dev(handle)

    def close(self):
        if self.h is not None:
            h, self.h = self.h, None
            self.engine._recycle_batch(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_engine = None
_engine_lock = threading.Lock()


def get_engine():
    """Per-process engine; a forked child gets a fresh one (never reuse a parent's CUDA context)."""
    global _engine
    with _engine_lock:
        if _engine is None or _engine.pid != os.getpid():
            _engine = Engine()
        return _engine
