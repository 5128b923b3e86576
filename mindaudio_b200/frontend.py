"""The conformer / deepspeech2 front-end (examples/conformer/dataset.py:117-168) on B200.

* ``compute_fbank_feats``  -- drop-in for the example's function (numpy in / numpy out).
* ``FbankPipeline``        -- the ragged, batched hot path: Kaldi-like 80-mel fbank (+ utterance
  CMVN, + global-CMVN statistics) over a flat waveform array and an offsets array, device-resident
  (torch tensors / raw device pointers) or host-to-host through pinned buffers.  This is what
  ``bench.py`` measures.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import _tables as T
from ._engine import get_engine


def conformer_plan(eng, sample_rate=16000, frame_len_ms=25, frame_shift_ms=10, mel_bin=80, n_fft=512, dither=0.0,
                   seed=0, allow_fast_path=True, low_freq=20.0, high_freq=8000.0, utt_cmvn_mean=False,
                   utt_cmvn_std=False):
    frame_len = sample_rate * frame_len_ms // 1000
    hop = sample_rate * frame_shift_ms // 1000
    if frame_len > n_fft:
        raise ValueError("frame length {} exceeds the {}-point FFT of the example (dataset.py:166)".format(frame_len, n_fft))
    # dataset.py:152: get_mel_banks(num_filter, 512, fs * 2, 20, 8000) with fs = sample_rate / 2
    bank = T.kaldi_triangle_bank(mel_bin, n_fft, float(sample_rate), low_freq, high_freq)
    return eng.plan(n_fft=n_fft, frame_len=frame_len, hop=hop, center=False, out_kind=L.OUT_LOGMEL,
                    window=T.povey_window(frame_len), preemph=0.97, remove_frame_mean=True, dither=dither,
                    dither_seed=seed, power=2.0, mel_fb=bank, log_kind=L.LOG_LN_EPS_IF_ZERO,
                    utt_cmvn_mean=utt_cmvn_mean, utt_cmvn_std=utt_cmvn_std, allow_fast_path=allow_fast_path)


def compute_fbank_feats(wav, sample_rate, frame_len, frame_shift, mel_bin, dither=0.0, seed=0, allow_fast_path=True):
    """``examples/conformer/dataset.py:159-168``: ``wav`` is the int16-scaled waveform
    (``read() * (1 << 15)``, ``:389-390``); returns float64 ``[T, mel_bin]`` log-mel energies.
    ``dither`` is an extension (the reference raises NotImplementedError, ``:557-558``)."""
    wav = np.asarray(wav)
    eng = get_engine()
    plan = conformer_plan(eng, sample_rate, frame_len, frame_shift, mel_bin, dither=dither, seed=seed,
                          allow_fast_path=allow_fast_path)
    x = np.ascontiguousarray(wav, dtype=np.float32).reshape(-1)
    out, _ = eng.run_frontend(plan, x, np.array([0, x.shape[0]], dtype=np.int64))
    return out.astype(np.float64)


def ds2_features(audio, n_fft=320, hop_length=160, win_length=320, window="hann", normalize=True):
    """The deepspeech2 front-end, ``examples/deepspeech2/dataset.py:36-47`` (``parse_audio``): ``stft`` (hann -- the
    YAML's hamming is never passed --, centred, constant padding) -> ``magphase(power=1)`` -> ``log1p`` -> scalar
    ``(m - m.mean()) / m.std()`` over the whole matrix.  One transform kernel writes ``log1p(|X|)`` directly (no complex
    spectrum, no separate magnitude pass) and accumulates the utterance's moments (plan flag ``utt_scalar_norm``); one
    streaming pass normalises.  float32 ``[n_fft // 2 + 1, T]`` like the reference.  ``audio``: a 1-D waveform, or a list
    of waveforms (one ragged batch, one launch; every utterance normalised by its own moments) -> a list of matrices."""
    many = isinstance(audio, (list, tuple))
    xs = [np.ascontiguousarray(np.asarray(a), dtype=np.float32).reshape(-1) for a in (audio if many else [audio])]
    for x in xs:
        if n_fft > x.shape[0]:
            raise ValueError("n_fft={} is too small for input signal of length={}".format(n_fft, x.shape[0]))
    eng = get_engine()
    plan = eng.plan(n_fft=n_fft, hop=hop_length, center=True, pad_mode="constant", out_kind=L.OUT_POWER,
                    window=T.analysis_window(window, win_length, n_fft), power=1.0, log_kind=L.LOG_LN_PLUS, log_arg=1.0,
                    utt_scalar_norm=bool(normalize))
    offs = np.zeros(len(xs) + 1, dtype=np.int64)
    np.cumsum([x.shape[0] for x in xs], out=offs[1:])
    flat = xs[0] if len(xs) == 1 else np.concatenate(xs)
    with eng.lock:
        b = eng.batch(plan, offs)
        try:
            out = np.empty((b.total_frames, plan.out_dim), dtype=np.float32)
            fo = np.array(b.frame_offsets, dtype=np.int64)
            dw = eng.buf("wave", flat.nbytes)
            do = eng.buf("out", out.nbytes)
            keep = eng.h2d(dw, flat)
            L.check(eng.lib.mafe_frontend_run(eng.ctx, plan.h, b.h, dw, L.WAVE_F32, 1.0, do, L.DBGROUP_NONE))
            eng.d2h(out, do)
            eng.sync()
            del keep
        finally:
            b.close()
    mats = [np.ascontiguousarray(out[fo[i]:fo[i + 1]].T) for i in range(len(xs))]
    return mats if many else mats[0]


class FbankPipeline:
    """Ragged batched fbank (+CMVN) on one GPU.

    ``layout(lengths)`` builds the device tables of a batch (sample/frame offsets, tile table);
    ``run(...)`` enqueues the kernels on the engine's stream (set it to torch's current stream
    with ``use_torch_stream()``); outputs are flat, frame-major ``[sum T_u, mel_bin]`` float32.
    """

    def __init__(self, sample_rate=16000, frame_len_ms=25, frame_shift_ms=10, mel_bin=80, n_fft=512, dither=0.0,
                 seed=0, cmvn="utt", mean_norm=True, std_norm=True, allow_fast_path=True, engine=None):
        self.eng = engine or get_engine()
        assert cmvn in (None, "utt")
        self.cmvn, self.mean_norm, self.std_norm = cmvn, mean_norm, std_norm
        fused = cmvn == "utt"
        self.plan = conformer_plan(self.eng, sample_rate, frame_len_ms, frame_shift_ms, mel_bin, n_fft, dither, seed,
                                   allow_fast_path, utt_cmvn_mean=fused and mean_norm, utt_cmvn_std=fused and std_norm)
        # the un-normalised plan serves the global-CMVN statistics pass (compute_cmvn_stats.py works on raw fbank)
        self.raw_plan = self.plan if not fused else conformer_plan(self.eng, sample_rate, frame_len_ms, frame_shift_ms,
                                                                   mel_bin, n_fft, dither, seed, allow_fast_path)
        self.mel_bin = mel_bin
        self.sample_rate = sample_rate

    # ---- layout ----
    def layout(self, lengths):
        so = np.zeros(len(lengths) + 1, dtype=np.int64)
        np.cumsum(np.asarray(lengths, dtype=np.int64), out=so[1:])
        return self.eng.batch(self.plan, so)

    def use_torch_stream(self):
        import torch
        # torch's default stream has handle 0; pass cudaStreamLegacy (0x1) so it is not mistaken for "own stream"
        self.eng.set_stream(torch.cuda.current_stream().cuda_stream or 1)

    # ---- device-resident ----
    def run(self, wave_ptr, batch, out_ptr, wave_dtype=L.WAVE_F32, wave_scale=1.0):
        """wave_ptr / out_ptr: device pointers (ints); fbank (+ fused utterance CMVN) of the whole batch."""
        eng, lib = self.eng, self.eng.lib
        L.check(lib.mafe_frontend_run(eng.ctx, self.plan.h, batch.h, C.c_void_p(wave_ptr), wave_dtype,
                                      float(wave_scale), C.c_void_p(out_ptr), L.DBGROUP_NONE))

    def run_global_stats(self, wave_ptr, batch, out_ptr, stats_ptr, wave_dtype=L.WAVE_F32, wave_scale=1.0):
        """Raw fbank into out_ptr and its global-CMVN sufficient statistics (sum x, sum x^2, N) ACCUMULATED into
        the device double[2*D+1] at stats_ptr (examples/conformer/compute_cmvn_stats.py:45-65, 104-112)."""
        eng, lib = self.eng, self.eng.lib
        L.check(lib.mafe_frontend_run(eng.ctx, self.raw_plan.h, batch.h, C.c_void_p(wave_ptr), wave_dtype,
                                      float(wave_scale), C.c_void_p(out_ptr), L.DBGROUP_NONE))
        L.check(lib.mafe_cmvn_stats_accumulate(eng.ctx, C.c_void_p(out_ptr), batch.total_frames, self.mel_bin,
                                               C.c_void_p(stats_ptr)))

    def __call__(self, wave, lengths=None, batch=None, out=None, wave_scale=1.0):
        """torch front door: ``wave`` flat CUDA tensor (float32 or int16)."""
        import torch
        own = batch is None
        if own:
            batch = self.layout(lengths)
        try:
            self.use_torch_stream()
            if out is None:
                out = torch.empty((batch.total_frames, self.mel_bin), dtype=torch.float32, device=wave.device)
            dt = L.WAVE_I16 if wave.dtype == torch.int16 else L.WAVE_F32
            self.run(wave.data_ptr(), batch, out.data_ptr(), dt, wave_scale)
            if own:
                torch.cuda.current_stream().synchronize()
            return out
        finally:
            if own:
                batch.close()

    # ---- host to host (the end-to-end path) ----
    def run_host(self, wave_host_ptr, sample_offsets, out_host_ptr, wave_dtype=L.WAVE_F32, wave_scale=1.0, chunk_utts=512):
        """Waveforms in (pinned) host memory -> features in host memory through ``mafe_frontend_run_host``:
        chunks of ``chunk_utts`` utterances on three streams, so H2D, kernels and D2H overlap.  Synchronous.
        ``sample_offsets``: int64 ``[n_utts + 1]`` into the flat host waveform.  Returns the frame offsets."""
        so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
        fo = np.empty(len(so), dtype=np.int64)
        eng = self.eng
        L.check(eng.lib.mafe_frontend_run_host(eng.ctx, self.plan.h, so.ctypes.data_as(C.c_void_p), len(so) - 1,
                                               C.c_void_p(wave_host_ptr), wave_dtype, float(wave_scale),
                                               C.c_void_p(out_host_ptr), fo.ctypes.data_as(C.c_void_p), int(chunk_utts),
                                               L.DBGROUP_NONE))
        return fo

    def features(self, waves, wave_scale=1.0, chunk_utts=512):
        """numpy front door: list of 1-D waveforms (float32 or int16) -> (flat features, frame offsets)."""
        lens = [len(w) for w in waves]
        dt = np.int16 if all(np.asarray(w).dtype == np.int16 for w in waves) else np.float32
        flat = np.ascontiguousarray(np.concatenate([np.asarray(w, dtype=dt) for w in waves])) if waves else np.zeros(0, dt)
        so = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=so[1:])
        total = sum(self.plan.num_frames(n) for n in lens)
        out = np.empty((total, self.mel_bin), dtype=np.float32)
        fo = self.run_host(flat.ctypes.data, so, out.ctypes.data, L.WAVE_I16 if dt == np.int16 else L.WAVE_F32,
                           wave_scale, chunk_utts)
        return out, fo

    def features_padded(self, waves, max_len=None, padding_value=0.0, wave_scale=1.0, spec_aug_conf=None, rng=None,
                        sort_by_length=False):
        """Front-end + collate in one device round trip (examples/conformer/dataset.py:456-491, 563-569, 616-621):
        list of 1-D waveforms -> ``(xs_pad [B, max_len, mel_bin] float32, xs_lengths [B] int32, xs_masks [B, 1, max_len]
        float32)``.  The ragged feature matrix never leaves the GPU: the padded batch and the mask are written by
        ``mafe_pad_sequence`` and only they are copied back.  Utterances longer than ``max_len`` are truncated like
        ``pad_sequence`` does; ``xs_lengths`` are the untruncated frame counts, as in the reference.
        ``spec_aug_conf`` (dataset.py:493-534) masks time / frequency rectangles of the ragged features on the device
        before the padding; the positions come from ``rng`` (a ``random.Random``; default the ``random`` module) with
        the reference's call sequence.

        ``sort_by_length=True`` reproduces the reference collate to the letter: ``extract_feature`` sorts the utterances by
        frame count, longest first (``np.argsort(lengths)[::-1]``, dataset.py:483-489), BEFORE ``spec_aug`` and
        ``pad_sequence`` -- so the random draws land on the utterances in that order and the batch rows come out in it.
        A fourth value, ``order`` (row i = input utterance ``order[i]``), is then returned.  With the default ``False`` rows
        and draws follow the input order (callers that pre-sort get the reference's result either way)."""
        dt = np.int16 if waves and all(np.asarray(w).dtype == np.int16 for w in waves) else np.float32
        arrs = [np.ascontiguousarray(np.asarray(w), dtype=dt).reshape(-1) for w in waves]   # no copy when already dt
        so = np.zeros(len(arrs) + 1, dtype=np.int64)
        np.cumsum([a.shape[0] for a in arrs], out=so[1:])
        eng = self.eng
        with eng.lock:
            d_w = eng.buf("wave", max(int(so[-1]) * np.dtype(dt).itemsize, 16))
            eng.h2d_gather(d_w, arrs)       # gathered into pinned staging by the library's host threads, one upload
            res = self._padded_from_device(d_w, L.WAVE_I16 if dt == np.int16 else L.WAVE_F32, so, max_len, padding_value,
                                           wave_scale, spec_aug_conf, rng, sort_by_length)
        return res

    def features_from_wav(self, files, max_len=None, padding_value=0.0, spec_aug_conf=None, rng=None, speeds=None,
                          sort_by_length=False):
        """WAV files -> padded feature batch, the whole ``read -> * (1 << 15) -> compute_fbank_feats -> CMVN ->
        spec_aug -> pad_sequence`` chain of the conformer input pipeline (examples/conformer/dataset.py:384-395,
        456-534, 563-621) with one upload (the files' PCM payloads, 2 bytes / sample for PCM16) and one download (the
        padded batch + mask).  ``files``: paths or binary file objects of mono WAV files
        (:func:`mindaudio_b200.data.io.load_batch`); ``speeds``: optional per-file speed-perturbation factors
        (dataset.py:391-404, Fourier resampling on the device).  Returns ``(xs_pad, xs_lengths, xs_masks)`` like
        :meth:`features_padded` (``sort_by_length=True``: the reference's longest-first order, plus ``order``)."""
        from .data.io import load_batch
        eng = self.eng
        with eng.lock:
            wb = load_batch(files, int16_scaled=True, speeds=speeds)
            for sr in wb.sample_rates:
                if sr != self.sample_rate:
                    raise ValueError("features_from_wav: file sampled at %d Hz, pipeline built for %d Hz" % (sr, self.sample_rate))
            return self._padded_from_device(wb.wave_dev, wb.dtype, wb.sample_offsets, max_len, padding_value, wb.wave_scale,
                                            spec_aug_conf, rng, sort_by_length)

    def _padded_from_device(self, d_w, wave_dtype, so, max_len, padding_value, wave_scale, spec_aug_conf, rng,
                            sort_by_length=False):
        """front-end + spec_aug + pad_sequence on a device-resident flat waveform batch (engine lock held)."""
        eng = self.eng
        n = len(so) - 1
        with eng.lock:
            b = eng.batch(self.plan, so)
            try:
                xs_lengths = np.diff(b.frame_offsets).astype(np.int32)
                # the reference's batch order (dataset.py:483-489); the device works in input order, the draws of spec_aug
                # are dealt in `order` and the rows are permuted on the way out
                order = np.argsort(xs_lengths)[::-1] if sort_by_length else None
                if max_len is None:
                    max_len = int(xs_lengths.max()) if n else 0
                xs_pad = np.empty((n, max_len, self.mel_bin), dtype=np.float32)
                xs_masks = np.empty((n, 1, max_len), dtype=np.float32)
                if xs_pad.size:
                    d_f = eng.buf("out", max(b.total_frames * self.mel_bin * 4, 16))
                    d_p = eng.buf("pad", xs_pad.nbytes)
                    d_m = eng.buf("aux", xs_masks.nbytes)
                    self.run(d_w.value if hasattr(d_w, "value") else d_w, b, d_f.value if hasattr(d_f, "value") else d_f,
                             wave_dtype, wave_scale)
                    keep_r = None
                    if spec_aug_conf:
                        from .data.masking import spec_aug_rects
                        seq = xs_lengths if order is None else xs_lengths[order]
                        rects = spec_aug_rects([(int(t), self.mel_bin) for t in seq], spec_aug_conf, rng)
                        if order is not None and len(rects):
                            rects[:, 0] = order[rects[:, 0]].astype(rects.dtype)
                        if len(rects):
                            d_r = eng.buf("rects", rects.nbytes)
                            keep_r = eng.h2d(d_r, rects)
                            L.check(eng.lib.mafe_mask_rects(eng.ctx, d_f, C.c_void_p(b.frame_offsets_dev), n, self.mel_bin,
                                                            d_r, len(rects), 0.0))
                    L.check(eng.lib.mafe_pad_sequence(eng.ctx, d_f, C.c_void_p(b.frame_offsets_dev), n, self.mel_bin,
                                                      max_len, float(padding_value), 1, d_p, d_m))
                    eng.d2h(xs_masks, d_m)
                    eng.d2h_staged(xs_pad, d_p)      # the big one: pinned staging + multi-threaded copy-out (synchronises)
                    eng.sync()
                    del keep_r
                if order is not None:
                    return xs_pad[order], xs_lengths[order], xs_masks[order], order
                return xs_pad, xs_lengths, xs_masks
            finally:
                b.close()

