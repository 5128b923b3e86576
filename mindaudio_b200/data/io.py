"""WAV input for the feature path ("next" row f3 of the scope table): ``mindaudio/data/io.py:541-747`` (``read``).

* the RIFF / RIFX container walk is native host code (``mafe_wav_parse``: the reference's chunk state machine over
  a byte buffer, quirks included -- ``offset`` skips BYTES (io.py:489-491), ``duration`` counts items over all
  channels (io.py:497-498), ``raise <str>`` for a non-WAVE form type surfaces as TypeError (io.py:676));
* the sample arithmetic of the "unified output format" (io.py:741-746: int16 / 32768, int32 and left-justified
  24-bit / 2^31) runs in ``mafe_wav_decode`` on the GPU, for ``read`` (numpy out, like every call of this package)
  and for :func:`load_batch` (payloads of many files -> one device-resident waveform batch for the front-end).

There is no CPU decode path: integer PCM goes through the device kernel, float / 8-bit payloads are returned as the
dtype view of the file's bytes exactly as the reference returns them (no arithmetic is defined for them).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import warnings

import numpy as np

from .. import _lib as L
from .._engine import get_engine

__all__ = ["read", "wav_info", "WavFileWarning", "load_batch", "WavBatch", "resampled_length"]


class WavFileWarning(UserWarning):
    """io.py:339-340."""


_ERRORS = {L.WAV_ERR_VALUE: ValueError, L.WAV_ERR_TYPE: TypeError, L.WAV_ERR_UNBOUND: UnboundLocalError,
           L.WAV_ERR_ZERODIV: ZeroDivisionError, L.WAV_ERR_STRUCT: struct.error, L.WAV_ERR_OS: OSError}
_DTYPES = {L.WAV_U8: "u1", L.WAV_I8: "i1", L.WAV_I16: "i2", L.WAV_I32: "i4", L.WAV_I64: "i8", L.WAV_F32: "f4", L.WAV_F64: "f8"}


def wav_info(raw, offset=0.0, duration=None, filelike=False):
    """Walk the container held in ``raw`` (bytes-like) the way ``read(file, offset, duration)`` does.  Returns the
    filled :class:`mindaudio_b200._lib.WavInfo`; raises what the reference raises; warns what it warns."""
    buf = np.frombuffer(raw, dtype=np.uint8)
    info = L.WavInfo()
    lib = L.load()
    rc = lib.mafe_wav_parse(buf.ctypes.data_as(C.c_void_p), buf.size, float(offset or 0.0), float(duration or 0.0),
                            int(bool(filelike)), C.byref(info))
    if rc != L.OK:
        msg = lib.mafe_last_error().decode("utf-8", "replace")
        raise _ERRORS.get(info.error_kind, L.MafeError)(msg)
    if info.warnings & L.WAV_WARN_UNKNOWN_CHUNK:
        warnings.warn("Chunk (non-data) not understood, skipping it.", WavFileWarning, stacklevel=3)
    if info.warnings & L.WAV_WARN_INCOMPLETE_ID:
        warnings.warn("Incomplete chunk ID, ignoring it.", WavFileWarning, stacklevel=3)
    if info.warnings & L.WAV_WARN_EOF:
        warnings.warn("Reached EOF prematurely; finished at {:d} bytes.".format(buf.size), WavFileWarning, stacklevel=3)
    return info


def _file_bytes(file):
    """(bytes of the file from its current position, filelike flag) -- io.py:644-647, 500-503."""
    if isinstance(file, (bytes, bytearray, memoryview)):
        return file, False                    # the contents of a file, already read
    if hasattr(file, "read"):
        try:
            file.fileno()
            filelike = False
        except (OSError, AttributeError):   # io.UnsupportedOperation is an OSError: np.fromfile cannot be used (io.py:500)
            filelike = True
        try:
            raw = file.read()
        finally:
            file.seek(0)                      # io.py:738-739
        return raw, filelike
    with open(file, "rb") as fh:
        return fh.read(), False


def _decode_device(raw_u8, info, out_dtype, scale=1.0):
    """Integer PCM payload -> float array through ``mafe_wav_decode`` (numpy in / numpy out)."""
    n = int(info.n_items)
    bps = {L.WAV_I16: 2, L.WAV_I24: 3, L.WAV_I32: 4}[info.sample_kind]
    out = np.empty(n, dtype=out_dtype)
    if n:
        payload = raw_u8[info.data_offset: info.data_offset + n * bps]
        eng = get_engine()
        with eng.lock:
            d_in = eng.buf("wave", max(payload.nbytes, 16))
            d_out = eng.buf("out", out.nbytes)
            keep = eng.h2d(d_in, payload)
            L.check(eng.lib.mafe_wav_decode(eng.ctx, d_in, n, info.sample_kind, info.big_endian,
                                            L.WAV_OUT_F64 if out.dtype == np.float64 else L.WAV_OUT_F32, float(scale), d_out))
            eng.d2h(out, d_out)
            eng.sync()
            del keep
    return out


def read(file, offset=0.0, duration=None):
    """``mindaudio.data.io.read`` (io.py:552-747): ``(audio, samplerate)``.

    Little-endian int16 PCM -> float64 in [-1, 1) (``/ 32768``), int32 and 24-bit PCM -> float64 (``/ 2147483648``);
    8-bit PCM -> uint8 unchanged, IEEE float -> the file's float dtype (byte order included), 5/6/7-byte containers ->
    int64 left-justified, RIFX (big-endian) integer PCM -> the big-endian integers unscaled (the reference's dtype
    test at io.py:741-746 never matches a big-endian dtype); 1-D for one channel, ``(n, channels)`` otherwise."""
    raw, filelike = _file_bytes(file)
    info = wav_info(raw, offset, duration, filelike)
    u8 = np.frombuffer(raw, dtype=np.uint8)
    kind, n, e = info.sample_kind, int(info.n_items), (">" if info.big_endian else "<")
    if kind in (L.WAV_I16, L.WAV_I24, L.WAV_I32) and not info.big_endian:
        audio = _decode_device(u8, info, np.float64)
    elif kind in (L.WAV_I24, L.WAV_I40, L.WAV_I48, L.WAV_I56):
        # byte placement only (io.py:505-512).  No arithmetic is applied to int64 containers, nor to RIFX integer
        # files: the reference compares the dtype with "int16" / "int32", which a big-endian dtype never equals
        # (io.py:741-746), so it returns the left-justified integers as they are.
        bps, wide = info.bytes_per_sample, (4 if kind == L.WAV_I24 else 8)
        a = np.zeros((n, wide), dtype=np.uint8)
        rows = u8[info.data_offset: info.data_offset + n * bps].reshape(n, bps)
        if info.big_endian:
            a[:, :bps] = rows
        else:
            a[:, wide - bps:] = rows
        audio = a.view(e + "i%d" % wide).reshape(n)
    else:
        audio = np.array(np.frombuffer(raw, dtype=e + _DTYPES[kind] if kind not in (L.WAV_U8, L.WAV_I8) else _DTYPES[kind],
                                       count=n, offset=int(info.data_offset)))
    if info.channels > 1:
        audio = audio.reshape(-1, info.channels)
    return audio, int(np.uint32(info.sample_rate))


class WavBatch:
    """A batch of decoded mono waveforms resident on the GPU: ``wave_dev`` (device pointer), ``dtype`` (``L.WAVE_I16``
    or ``L.WAVE_F32``), ``sample_offsets`` int64 ``[B + 1]``, ``sample_rates``."""

    def __init__(self, wave_dev, dtype, sample_offsets, sample_rates, wave_scale):
        self.wave_dev, self.dtype, self.sample_offsets = wave_dev, dtype, sample_offsets
        self.sample_rates, self.wave_scale = sample_rates, wave_scale

    @property
    def lengths(self):
        return np.diff(self.sample_offsets)


def _name(f):
    return "<%d bytes>" % len(f) if isinstance(f, (bytes, bytearray, memoryview)) else f


class _BatchWalk:
    """``mafe_wav_stage`` over the contents of many files: ``walk()`` parses all containers on the library's host
    threads (path-like semantics, offset 0, no duration), ``pack(ptr, nbytes)`` copies their payloads back to back into
    a pinned buffer.  ``infos[k]`` / ``offsets`` (payload byte offsets, ``[n + 1]``) are valid after ``walk()``."""

    def __init__(self, raws, files):
        self.n, self.files = len(raws), files
        self._views = [np.frombuffer(r, dtype=np.uint8) for r in raws]          # keeps the buffers alive
        self._ptrs = (C.c_void_p * max(self.n, 1))(*[v.ctypes.data for v in self._views])
        self._sizes = np.array([v.size for v in self._views], dtype=np.int64)
        self._infos = (L.WavInfo * max(self.n, 1))()
        self.offsets = np.zeros(self.n + 1, dtype=np.int64)
        self.infos = []

    def _call(self, stage_ptr, stage_bytes):
        lib = L.load()
        failed = C.c_int32(-1)
        rc = lib.mafe_wav_stage(self._ptrs, self._sizes.ctypes.data_as(C.c_void_p), self.n, 0, self._infos,
                                self.offsets.ctypes.data_as(C.c_void_p), stage_ptr, int(stage_bytes), C.byref(failed))
        if rc != L.OK:
            msg = lib.mafe_last_error().decode("utf-8", "replace")
            if failed.value >= 0:
                exc = _ERRORS.get(self._infos[failed.value].error_kind, L.MafeError)
                raise exc("%s (file %r)" % (msg, _name(self.files[failed.value])))
            L.check(rc)

    def walk(self):
        self._call(None, 0)
        self.infos = [self._infos[k] for k in range(self.n)]
        for k, i in enumerate(self.infos):
            if i.warnings:
                warnings.warn("%r: WAV container with unknown chunks or a short data chunk" % (_name(self.files[k]),),
                              WavFileWarning, stacklevel=4)
        return self

    def pack(self, stage_ptr, stage_bytes):
        self._call(stage_ptr, stage_bytes)

    def close(self):
        pass


class _FileWalk:
    """``mafe_wav_files_*``: the same as :class:`_BatchWalk` for paths -- the library's host threads map the files,
    walk the containers and later copy the payloads from the page cache straight into the pinned staging buffer (no
    Python-side ``read()``, no intermediate copy)."""

    def __init__(self, paths):
        self.n, self.files = len(paths), list(paths)
        enc = [os.fsencode(p) for p in self.files]
        self._paths = (C.c_char_p * max(self.n, 1))(*enc)
        self._infos = (L.WavInfo * max(self.n, 1))()
        self.offsets = np.zeros(self.n + 1, dtype=np.int64)
        self.infos = []
        self._h = C.c_void_p()

    def walk(self):
        lib = L.load()
        failed = C.c_int32(-1)
        rc = lib.mafe_wav_files_open(self._paths, self.n, 0, C.byref(self._h), self._infos,
                                     self.offsets.ctypes.data_as(C.c_void_p), C.byref(failed))
        if rc != L.OK:
            msg = lib.mafe_last_error().decode("utf-8", "replace")
            if failed.value >= 0:
                kind = self._infos[failed.value].error_kind
                if kind == L.WAV_ERR_OS:
                    raise OSError(msg)
                raise _ERRORS.get(kind, L.MafeError)("%s (file %r)" % (msg, self.files[failed.value]))
            L.check(rc)
        self.infos = [self._infos[k] for k in range(self.n)]
        for k, i in enumerate(self.infos):
            if i.warnings:
                warnings.warn("%r: WAV container with unknown chunks or a short data chunk" % (self.files[k],),
                              WavFileWarning, stacklevel=4)
        return self

    def pack(self, stage_ptr, stage_bytes):
        L.check(L.load().mafe_wav_files_pack(self._h, stage_ptr, int(stage_bytes)))

    def close(self):
        if self._h:
            L.load().mafe_wav_files_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001 -- interpreter shutdown
            pass


def resampled_length(n_in, orig_freq, new_freq):
    """Output length of ``processing.resample`` (processing.py:170-172), with its float arithmetic."""
    ratio = float(new_freq) / orig_freq
    return int(np.ceil(n_in * ratio))


def load_batch(files, int16_scaled=True, buffer_name="wavbatch", speeds=None):
    """Decode mono WAV files straight into one device-resident waveform batch (no host-side float pass).

    ``int16_scaled=True`` produces what the conformer pipeline feeds its front-end, ``read(path) * (1 << 15)``
    (examples/conformer/dataset.py:389-390): PCM16 files are uploaded as they are (2 bytes / sample) and consumed by
    the front-end's int16 input path; every other format is decoded on the device to float32 with the factor folded
    into the decode.  ``int16_scaled=False`` gives ``read(path)`` as float32.

    ``speeds``: optional per-file speed factors -- the conformer pipeline's speed perturbation
    (examples/conformer/dataset.py:391-404: ``resample(waveform, sample_rate * speed, sample_rate)`` for
    ``speed != 1.0``, the Fourier method of ``processing.resample``), run on the device in float64 between the
    decode and the float32 batch; the caller draws the factors (the reference: ``random.choice([0.9, 1.0, 1.1])``).

    Returns a :class:`WavBatch`; its device buffer belongs to the engine (name ``buffer_name``) and stays valid until
    the next ``load_batch`` with the same name."""
    files = list(files)
    if files and all(isinstance(f, (str, os.PathLike)) for f in files):
        walk = _FileWalk(files).walk()              # paths: mapped, walked and packed by the library's host threads
    else:
        walk = _BatchWalk([_file_bytes(f)[0] for f in files], files).walk()
    infos, bo = walk.infos, walk.offsets
    for f, info in zip(files, infos):
        if info.channels != 1:
            raise ValueError("load_batch: %r has %d channels; the feature front-end takes mono waveforms" % (_name(f), info.channels))
        if info.sample_kind in (L.WAV_I40, L.WAV_I48, L.WAV_I56, L.WAV_I64):
            raise ValueError("load_batch: %r holds 64-bit integer samples, which have no unit-range scaling in read()" % (_name(f),))
    n = len(infos)
    rates = [int(np.uint32(i.sample_rate)) for i in infos]
    if speeds is None:
        speeds = [1.0] * n
    speeds = [float(v) for v in speeds]
    if len(speeds) != n:
        raise ValueError("load_batch: %d speed factors for %d files" % (len(speeds), n))
    if any(v <= 0 for v in speeds):
        raise ValueError("load_batch: speed factors must be positive")
    n_in = [int(i.n_items) for i in infos]
    n_out = [m if v == 1.0 or m == 0 else resampled_length(m, rates[k] * v, rates[k]) for k, (m, v) in enumerate(zip(n_in, speeds))]
    so = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(n_out, out=so[1:])
    total = int(so[-1])
    scale = 32768.0 if int16_scaled else 1.0
    perturbed = [v != 1.0 and m > 0 for v, m in zip(speeds, n_in)]
    all_pcm16 = (int16_scaled and n > 0 and not any(perturbed)
                 and all(i.sample_kind == L.WAV_I16 and not i.big_endian for i in infos))
    eng = get_engine()
    item_out = 2 if all_pcm16 else 4

    def factor(i):
        # read() leaves RIFX integer PCM unscaled (io.py:741-746 never matches a big-endian dtype)
        undo = 1.0
        if i.big_endian:
            undo = {L.WAV_I16: 32768.0, L.WAV_I24: 2147483648.0, L.WAV_I32: 2147483648.0}.get(i.sample_kind, 1.0)
        return scale * undo

    with eng.lock:
        d_wave = eng.buf(buffer_name, max(total * item_out, 16))
        # one pinned staging buffer for all payloads, packed by the library's host threads: a single H2D
        n_stage = int(bo[-1])
        stage_ptr, _ = eng.pinned("wavstage", max(n_stage, 16))
        walk.pack(stage_ptr, n_stage)
        walk.close()                                  # the files' mappings are no longer needed
        if all_pcm16:
            eng.h2d_raw(d_wave, stage_ptr, n_stage)                                  # the payload IS the int16 batch
        else:
            d_stage = eng.buf("wavstage", max(n_stage, 16))
            eng.h2d_raw(d_stage, stage_ptr, n_stage)
            k = 0
            while k < n:
                src = C.c_void_p(d_stage.value + int(bo[k]))
                dst = C.c_void_p(d_wave.value + int(so[k]) * item_out)
                if perturbed[k]:
                    # decode -> float64, Fourier resampling in float64 (mafe_resample_fft), -> float32 into the batch
                    need = C.c_size_t()
                    L.check(eng.lib.mafe_resample_workspace(1, n_in[k], n_out[k], C.byref(need)))
                    d_x, d_y = eng.buf("rs_in", n_in[k] * 8), eng.buf("rs_out", n_out[k] * 8)
                    d_w = eng.buf("work", need.value)
                    L.check(eng.lib.mafe_wav_decode(eng.ctx, src, n_in[k], infos[k].sample_kind, infos[k].big_endian,
                                                    L.WAV_OUT_F64, factor(infos[k]), d_x))
                    L.check(eng.lib.mafe_resample_fft(eng.ctx, d_x, 1, n_in[k], n_out[k], d_y, d_w, need.value))
                    L.check(eng.lib.mafe_wav_decode(eng.ctx, d_y, n_out[k], L.WAV_F64, 0, L.WAV_OUT_F32, 1.0, dst))
                    k += 1
                    continue
                j = k
                while (j + 1 < n and not perturbed[j + 1]
                       and (infos[j + 1].sample_kind, infos[j + 1].big_endian) == (infos[k].sample_kind, infos[k].big_endian)):
                    j += 1
                items = int(so[j + 1] - so[k])
                if items:
                    L.check(eng.lib.mafe_wav_decode(eng.ctx, src, items, infos[k].sample_kind, infos[k].big_endian,
                                                    L.WAV_OUT_F32, factor(infos[k]), dst))
                k = j + 1
        eng.sync()
    return WavBatch(d_wave, L.WAVE_I16 if all_pcm16 else L.WAVE_F32, so, rates, 1.0)

