"""Drop-in for ``mindaudio.data.features`` (mindaudio/data/features.py) on B200: ``fbank`` (alias
``fbanks`` -- README.md:41 / features.py:247 use that name), ``mfcc``, ``compute_deltas``,
``context_window``.  Same signatures, defaults and numpy semantics as the reference; arithmetic
in libmafe.so.  ``mfcc`` keeps the reference defaults ``deltas=True, context=True``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L
from .. import _tables as T
from .._engine import get_engine
from .._enums import BorderType, NormMode, WindowType
from . import spectrum as _sp

__all__ = [
    "context_window",
    "compute_deltas",
    "fbank",
    "fbanks",
    "mfcc",
    "spectral_centroid",
    "soft_mask",
    "hpss",
    "harmonic",
]


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _context_dev(eng, d_in, d_out, n_mats, f, t, left, right):
    L.check(eng.lib.mafe_context_window(eng.ctx, d_in, d_out, n_mats, f, t, left, right))


def context_window(waveforms, left_frames=0, right_frames=0):
    """``features.py:69-155``: gather +-k frames into one feature vector (the reference's grouped
    identity-kernel Conv1d, as a pure gather; float32 out).  ``[freq, time]``, ``[batch, freq,
    time]`` or ``[batch, channel, freq, time]``."""
    waveforms = np.asarray(waveforms)
    nd = waveforms.ndim
    if nd == 2:
        x = waveforms[None]
    elif nd == 3:
        x = waveforms
    elif nd == 4:
        # the reference moves channels last, folds (batch, time) and convolves along the channel
        # axis (features.py:108-126); reproduced literally
        b, ch, f, t = waveforms.shape
        x = waveforms.transpose((0, 2, 3, 1)).reshape((b * t, f, ch))
    else:
        raise TypeError("Input dimension must be 2, 3 or 4, but got {}".format(nd))
    x = np.ascontiguousarray(x, dtype=np.float32)
    n_mats, f, t = x.shape
    csize = left_frames + right_frames + 1
    out = np.empty((n_mats, f * csize, t), dtype=np.float32)
    if x.size:
        eng = get_engine()
        with eng.lock:
            dx = eng.buf("wave", x.nbytes)
            do = eng.buf("out", out.nbytes)
            keep = eng.h2d(dx, x)
            _context_dev(eng, dx, do, n_mats, f, t, left_frames, right_frames)
            eng.d2h(out, do)
            eng.sync()
            del keep
    # nd == 2: the reference tests `len(x_shape) == 2` on the EXPANDED shape (features.py:104-106,
    # 153-154), so the batch axis it added is never squeezed: [F, T] comes back as [1, F*C, T]
    if nd == 4:
        b, ch, f4, t4 = waveforms.shape
        out = out.reshape((b, out.shape[1], t4, out.shape[-1])).transpose((0, 3, 1, 2))
    return out


def compute_deltas(specgram, win_length=5, pad_mode="edge"):
    """``features.py:158-193`` (``msaudio.ComputeDeltas``): ``(..., freq, time)``."""
    specgram = np.asarray(specgram)
    pad_mode = BorderType(pad_mode)
    if win_length < 3:
        raise ValueError("win_length must be no less than 3, got {}".format(win_length))
    x = np.ascontiguousarray(specgram, dtype=np.float32)
    t = x.shape[-1]
    rows = x.size // t if t else 0
    out = np.empty_like(x)
    if x.size:
        eng = get_engine()
        with eng.lock:
            dx = eng.buf("wave", x.nbytes)
            do = eng.buf("out", out.nbytes)
            keep = eng.h2d(dx, x)
            L.check(eng.lib.mafe_compute_deltas(eng.ctx, dx, do, 1, rows, t, 0, 0, win_length, L.PAD[pad_mode.value]))
            eng.d2h(out, do)
            eng.sync()
            del keep
    return out.astype(np.float64 if specgram.dtype == np.float64 else np.float32, copy=False)


def _postprocess(eng, d_feat, n_mats, t, dim, deltas, context, left, right):
    """Device-side tail shared by fbank/mfcc: frame-major [n_mats*t, dim] -> [n_mats, D', t] with
    optional delta/delta-delta rows (features.py:264-267) and context window (:268-269)."""
    rows = dim * (3 if deltas else 1)
    d_a = eng.buf("post_a", 4 * n_mats * rows * t)
    L.check(eng.lib.mafe_transpose(eng.ctx, d_feat, d_a, n_mats, t, dim, rows * t))
    if deltas:
        edge = L.PAD["edge"]
        p1 = C.c_void_p(d_a.value + 4 * dim * t)
        p2 = C.c_void_p(d_a.value + 8 * dim * t)
        L.check(eng.lib.mafe_compute_deltas(eng.ctx, d_a, p1, n_mats, dim, t, rows * t, rows * t, 5, edge))
        L.check(eng.lib.mafe_compute_deltas(eng.ctx, p1, p2, n_mats, dim, t, rows * t, rows * t, 5, edge))
    if context:
        csize = left + right + 1
        d_b = eng.buf("post_b", 4 * n_mats * rows * csize * t)
        _context_dev(eng, d_a, d_b, n_mats, rows, t, left, right)
        return d_b, rows * csize
    return d_a, rows


def _features(waveforms, plan_kw, n_out, deltas, context, left_frames, right_frames, win_length, hop_length,
              window, n_fft, aux_mels=0):
    """``aux_mels`` > 0 (``mel_and_fbank`` / ``mel_and_mfcc``): also return the mel spectrogram ``[..., aux_mels, T]`` the
    features are the dB / DCT of, written by the same kernel launch (``mafe_frontend_run_aux``); kernels without the second
    output run a ``MAFE_OUT_MEL`` plan on the uploaded waveform instead."""
    waveforms = np.asarray(waveforms)
    if waveforms.ndim not in (1, 2, 3):
        raise TypeError("Unsupported MelSpectrogram shape {}".format(waveforms.ndim + 1))
    win_length = win_length if win_length is not None else n_fft
    hop_length = hop_length if hop_length is not None else win_length // 2
    eng = get_engine()
    plan = _sp._spectrogram_plan(eng, n_fft, win_length, hop_length, window, 2.0, False, True, "reflect", **plan_kw)
    x, lead = _sp._flatten_batch(waveforms)
    if x.shape[-1] < n_fft // 2 + 1:
        raise ValueError("padding of n_fft // 2 = {} needs a longer input than {}".format(n_fft // 2, x.shape[-1]))
    n_mats = x.shape[0]
    # amplitude_to_dB clamp group (spectrum.py:81-86): mel is [.., n_mels, T]
    #   1-D wave -> 2-D mel: the matrix; 2-D wave -> 3-D mel: the WHOLE batch; 3-D wave -> per batch item
    utt_group = None
    if plan_kw.get("log_kind") != L.LOG_DB:
        db_group = L.DBGROUP_NONE
    elif waveforms.ndim <= 2:
        db_group = L.DBGROUP_BATCH
    else:
        db_group = L.DBGROUP_MAP
        utt_group = np.repeat(np.arange(lead[0], dtype=np.int32), lead[1])
    with eng.lock:
        b = eng.batch(plan, _sp._dense_offsets(n_mats, x.shape[1]), utt_group)
        try:
            t = int(b.frame_offsets[1] - b.frame_offsets[0]) if n_mats else 0
            dw = eng.buf("wave", x.nbytes)
            do = eng.buf("out", 4 * max(b.total_frames, 1) * n_out)
            keep = eng.h2d(dw, x)
            mel = None
            if aux_mels:
                mel = np.empty((max(b.total_frames, 0), aux_mels), dtype=np.float32)
                dm = eng.buf("aux", 4 * max(b.total_frames, 1) * aux_mels)
                rc = eng.lib.mafe_frontend_run_aux(eng.ctx, plan.h, b.h, dw, L.WAVE_F32, 1.0, do, db_group, dm)
                if rc == L.E_UNSUPPORTED:     # this kernel has no second output: a MAFE_OUT_MEL plan beside the feature plan
                    mel_plan = _sp._spectrogram_plan(eng, n_fft, win_length, hop_length, window, 2.0, False, True, "reflect",
                                                     out_kind=L.OUT_MEL, mel_fb=plan_kw["mel_fb"])
                    b2 = eng.batch(mel_plan, _sp._dense_offsets(n_mats, x.shape[1]))
                    try:
                        L.check(eng.lib.mafe_frontend_run(eng.ctx, mel_plan.h, b2.h, dw, L.WAVE_F32, 1.0, dm, L.DBGROUP_NONE))
                    finally:
                        b2.close()
                    rc = eng.lib.mafe_frontend_run(eng.ctx, plan.h, b.h, dw, L.WAVE_F32, 1.0, do, db_group)
                L.check(rc)
                if mel.size:
                    eng.d2h(mel, dm)
            else:
                L.check(eng.lib.mafe_frontend_run(eng.ctx, plan.h, b.h, dw, L.WAVE_F32, 1.0, do, db_group))
            # [batch, channel, time] input: the reference's 4-D context route convolves along the channel
            # axis (features.py:108-126) -> done by context_window() on the finished array below
            ctx_dev = context and waveforms.ndim < 3
            d_res, rows = _postprocess(eng, do, n_mats, t, n_out, deltas, ctx_dev, left_frames, right_frames)
            # 2-D features through context_window come back [1, F*C, T] in the reference (see context_window)
            out = np.empty((((1,) if ctx_dev and not lead else lead)) + (rows, t), dtype=np.float32)
            eng.d2h(out, d_res)
            eng.sync()
            del keep
        finally:
            b.close()
    if context and not ctx_dev:
        out = context_window(out, left_frames, right_frames)
    out = out.astype(_sp._out_dtype(waveforms), copy=False)
    if aux_mels:
        mel = _sp._frames_to_ft(mel, lead, t, aux_mels).astype(_sp._out_dtype(waveforms), copy=False)
        return mel, out
    return out


def fbank(
    waveforms,
    deltas=False,
    context=False,
    n_mels=40,
    n_fft=400,
    sample_rate=16000,
    f_min=0.0,
    f_max=None,
    left_frames=5,
    right_frames=5,
    win_length=None,
    hop_length=None,
    window="hann",
):
    """``features.py:196-270``: ``amplitude_to_dB(melspectrogram(x), "power", ref=1.0, top_db=80)``
    with optional deltas (x3 rows) and context window; ``[..., n_mels', T]``."""
    bank = _sp._mel_bank(n_fft, n_mels, sample_rate, f_min, f_max, "none", "htk")
    kw = dict(out_kind=L.OUT_LOGMEL, mel_fb=bank, log_kind=L.LOG_DB, log_arg=1e-10, log_mult=10.0,
              log_offset=10.0 * np.log10(max(1e-10, 1.0)), top_db=80.0)
    out = _features(waveforms, kw, n_mels, deltas, context, left_frames, right_frames, win_length, hop_length,
                    window, n_fft)
    return out


def mel_and_fbank(waveforms, n_mels=40, n_fft=400, sample_rate=16000, f_min=0.0, f_max=None, win_length=None, hop_length=None,
                  window="hann"):
    """EXTENSION (not in the reference): ``(melspectrogram(x, ...), fbank(x, ...))`` of the same waveforms with the same front-end
    parameters from ONE transform -- what a pipeline that calls ``spectrum.melspectrogram`` (spectrum.py:609-698) next to
    ``features.fbank`` (features.py:196-270) computes twice.  The two arrays equal the separate calls' results
    (``melspectrogram(x, n_fft, win_length, hop_length, window=window, n_mels=n_mels, sample_rate=..., f_min=..., f_max=...)`` and
    ``fbank(x, n_mels=n_mels, ...)``): the n_fft 400 kernel writes the mel energies beside their dB (``mafe_frontend_run_aux``)."""
    bank = _sp._mel_bank(n_fft, n_mels, sample_rate, f_min, f_max, "none", "htk")
    kw = dict(out_kind=L.OUT_LOGMEL, mel_fb=bank, log_kind=L.LOG_DB, log_arg=1e-10, log_mult=10.0,
              log_offset=10.0 * np.log10(max(1e-10, 1.0)), top_db=80.0)
    return _features(waveforms, kw, n_mels, False, False, 0, 0, win_length, hop_length, window, n_fft, aux_mels=n_mels)


def mel_and_mfcc(waveforms, n_mels=23, n_mfcc=20, n_fft=400, sample_rate=16000, f_min=0.0, f_max=None, win_length=None,
                 hop_length=None, norm="ortho", log_mels=False):
    """EXTENSION (not in the reference): ``(melspectrogram(x, ...), mfcc(x, ..., deltas=False, context=False))`` from ONE transform
    (BASELINE configs[3]: the ECAPA-TDNN / fastspeech2 style front-end asks for both); see :func:`mel_and_fbank`."""
    norm = NormMode(norm)
    if n_mfcc > n_mels:
        raise ValueError("The number of MFCC coefficients must be no more than # mel bins.")
    dct = T.dct_matrix(n_mfcc, n_mels, norm)
    bank = _sp._mel_bank(n_fft, n_mels, sample_rate, f_min, f_max, "none", "htk")
    if log_mels:
        kw = dict(out_kind=L.OUT_MFCC, mel_fb=bank, dct=dct, log_kind=L.LOG_LN_PLUS, log_arg=1e-6)
    else:
        kw = dict(out_kind=L.OUT_MFCC, mel_fb=bank, dct=dct, log_kind=L.LOG_DB, log_arg=1e-10, log_mult=10.0,
                  log_offset=0.0, top_db=80.0)
    return _features(waveforms, kw, n_mfcc, False, False, 0, 0, win_length, hop_length, "hann", n_fft, aux_mels=n_mels)


fbanks = fbank  # README.md:41 and the docstring example (features.py:247) call it ``fbanks``


def mfcc(
    waveforms,
    deltas=True,
    context=True,
    n_mels=23,
    n_mfcc=20,
    n_fft=400,
    sample_rate=16000,
    f_min=0.0,
    f_max=None,
    left_frames=5,
    right_frames=5,
    win_length=None,
    hop_length=None,
    norm="ortho",
    log_mels=False,
):
    """``features.py:273-373``: DCT-II of the (dB or ``ln(mel + 1e-6)``) mel spectrogram;
    NOTE the reference defaults ``deltas=True, context=True``."""
    norm = NormMode(norm)
    if n_mfcc > n_mels:
        raise ValueError("The number of MFCC coefficients must be no more than # mel bins.")
    dct = T.dct_matrix(n_mfcc, n_mels, norm)
    bank = _sp._mel_bank(n_fft, n_mels, sample_rate, f_min, f_max, "none", "htk")
    if log_mels:
        kw = dict(out_kind=L.OUT_MFCC, mel_fb=bank, dct=dct, log_kind=L.LOG_LN_PLUS, log_arg=1e-6)
    else:
        kw = dict(out_kind=L.OUT_MFCC, mel_fb=bank, dct=dct, log_kind=L.LOG_DB, log_arg=1e-10, log_mult=10.0,
                  log_offset=0.0, top_db=80.0)
    out = _features(waveforms, kw, n_mfcc, deltas, context, left_frames, right_frames, win_length, hop_length,
                    "hann", n_fft)
    return out


def spectral_centroid(waveforms, sample_rate, n_fft=400, win_length=None, hop_length=None, pad=0, window="hann"):
    """``features.py:22-66`` (``msaudio.SpectralCentroid`` = torchaudio.functional.spectral_centroid): the
    magnitude-weighted mean frequency of every frame, ``[..., time]``.

    One front-end pass: the magnitude spectrogram (power 1, reflect-centred) is projected on the two-row "filterbank"
    ``[freqs; ones]`` inside the kernel, so only ``sum f|X|`` and ``sum |X|`` per frame leave the GPU."""
    from .spectrum import _frames_to_ft, _out_dtype, _run_spec, _spectrogram_plan
    waveforms = np.asarray(waveforms)
    win_length = win_length if win_length else n_fft
    hop_length = hop_length if hop_length else win_length // 2
    window = WindowType(window)
    n_bins = n_fft // 2 + 1
    bank = np.stack([np.linspace(0, sample_rate // 2, n_bins), np.ones(n_bins)]).astype(np.float32)
    plan = _spectrogram_plan(get_engine(), n_fft, win_length, hop_length, window, 1.0, False, True, "reflect",
                             out_kind=L.OUT_MEL, mel_fb=bank)
    out, lead, frames = _run_spec(plan, waveforms, pad, n_fft, True)
    num_den = _frames_to_ft(out, lead, frames, 2).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        cen = num_den[..., 0, :] / num_den[..., 1, :]
    return cen.astype(_out_dtype(waveforms), copy=False)


def _masks_device(harm, perc, margin_h, margin_p, power, split_zeros):
    harm = np.ascontiguousarray(harm, dtype=np.float32)
    perc = np.ascontiguousarray(perc, dtype=np.float32)
    mh, mp = np.empty_like(harm), np.empty_like(perc)
    if harm.size:
        eng = get_engine()
        with eng.lock:
            d_h, d_p = eng.buf("wave", harm.nbytes), eng.buf("out", perc.nbytes)
            d_mh, d_mp = eng.buf("pad", harm.nbytes), eng.buf("aux", perc.nbytes)
            keep = (eng.h2d(d_h, harm), eng.h2d(d_p, perc))
            L.check(eng.lib.mafe_hpss_masks(eng.ctx, d_h, d_p, harm.size, float(margin_h), float(margin_p), float(power),
                                            int(bool(split_zeros)), d_mh, d_mp))
            eng.d2h(mh, d_mh)
            eng.d2h(mp, d_mp)
            eng.sync()
            del keep
    return mh, mp


def soft_mask(x_input, x_ref, *, power=1, split_zeros=False):
    """``features.py:438-469``: ``x^p / (x^p + ref^p)`` computed on the scale ``max(x, ref)``; ``power=inf`` = hard mask."""
    x_input, x_ref = np.asarray(x_input), np.asarray(x_ref)
    if np.any(x_input < 0) or np.any(x_ref < 0):
        raise TypeError("x_input and x_ref must be non-negative")
    if x_input.shape != x_ref.shape:
        raise TypeError("x_input and x_ref shape mismatch.")
    if power <= 0:
        raise TypeError("power must be strictly positive.")
    mh, _ = _masks_device(x_input, x_ref, 1.0, 1.0, power, split_zeros)
    if not np.isfinite(power):
        return mh.astype(bool)
    return mh.astype(x_input.dtype if np.issubdtype(x_input.dtype, np.floating) else np.float32, copy=False)


def hpss(spectrogram, *, kernel_size=31, power=2.0, mask=False, margin=1.0):
    """``features.py:472-529``: median-filtering harmonic / percussive separation of a (complex or magnitude)
    spectrogram ``[..., F, T]``.  Magnitude, the two median filters (31 along time / frequency, scipy ``reflect``
    boundary) and the soft masks run on the GPU in one round trip."""
    spectrogram = np.asarray(spectrogram)
    if not np.iscomplexobj(spectrogram):
        phase = 1
        mag = spectrogram
    else:
        mag, phase = _sp.magphase(spectrogram, power=1)
    margin_harmonic, margin_perc = (margin[0], margin[1]) if not np.isscalar(margin) else (margin, margin)
    win_harmonic, win_perc = (kernel_size[0], kernel_size[1]) if not np.isscalar(kernel_size) else (kernel_size, kernel_size)
    if margin_harmonic < 1 or margin_perc < 1:
        raise TypeError("Margins must be >= 1.0. " "A typical range is between 1 and 10.")
    if power <= 0:
        raise TypeError("power must be strictly positive.")
    if mag.ndim < 2:
        raise ValueError("spectrogram must have at least 2 dimensions [..., F, T]")
    x = np.ascontiguousarray(mag, dtype=np.float32)
    F, T = x.shape[-2], x.shape[-1]
    n = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    split_zeros = margin_harmonic == 1 and margin_perc == 1
    mask_h, mask_p = np.empty_like(x), np.empty_like(x)
    if x.size:
        eng = get_engine()
        with eng.lock:
            d_x = eng.buf("wave", x.nbytes)
            d_h, d_p = eng.buf("out", x.nbytes), eng.buf("scratch", x.nbytes)
            d_mh, d_mp = eng.buf("pad", x.nbytes), eng.buf("aux", x.nbytes)
            keep = eng.h2d(d_x, x)
            L.check(eng.lib.mafe_median_filter(eng.ctx, d_x, d_h, n, F, T, int(win_harmonic), 1))   # along time
            L.check(eng.lib.mafe_median_filter(eng.ctx, d_x, d_p, n, F, T, int(win_perc), 0))       # along frequency
            L.check(eng.lib.mafe_hpss_masks(eng.ctx, d_h, d_p, x.size, float(margin_harmonic), float(margin_perc), float(power),
                                            int(split_zeros), d_mh, d_mp))
            eng.d2h(mask_h, d_mh)
            eng.d2h(mask_p, d_mp)
            eng.sync()
            del keep
    if not np.isfinite(power):
        mask_h, mask_p = mask_h.astype(bool), mask_p.astype(bool)
    elif mag.dtype == np.float64:
        mask_h, mask_p = mask_h.astype(np.float64), mask_p.astype(np.float64)
    if mask:
        return mask_h, mask_p
    return ((mag * mask_h) * phase, (mag * mask_p) * phase)


def harmonic(y_input, **kwargs):
    """``features.py:532-559``: ``istft(hpss(stft(y, n_fft=2048, pad_mode="constant"))[0], length=len(y))``."""
    y_input = np.asarray(y_input)
    y_stft = _sp.stft(y_input, n_fft=2048, pad_mode="constant")
    stft_harm = hpss(y_stft, **kwargs)[0]
    return _sp.istft(stft_harm, length=y_input.shape[-1])

