"""Freeze golden vectors by EXECUTING THE REFERENCE'S OWN PYTHON -- TEST INFRASTRUCTURE.

    python -m oracle.make_goldens            # dev container only (/root/reference must exist)

Writes ``tests/golden/*.npz``.  In-repo reference python (stft/istft/magphase/
dB, conformer fbank, CMVN pieces) runs unchanged via ``oracle.ref_loader``.
``spectrogram/melspectrogram/fbank/mfcc`` run the reference's python wrappers
around the restated mindspore ops (``oracle/ms_shim``) -- those entries carry
``msop=1`` and are "parity unpinned vs the MindSpore binary" (oracle/__init__.py);
``oracle/make_goldens_with_mindspore.py`` regenerates them where MindSpore 2.3.0 exists.

Inputs: the reference's sample WAV ``tests/samples/ASR/BAC009S0002W0122.wav`` (its
decoded int16 samples are stored in the fixture so tests do not need the
reference tree) and seeded synthetic waveforms.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(REPO, "tests", "golden")


def synth(seed, shape, scale=0.05):
    """Synthetic waveform of SURVEY.md section 8d: clip(0.05*N(0,1), -1, 1) float32."""
    rng = np.random.default_rng(seed)
    return np.clip(scale * rng.standard_normal(shape), -1.0, 1.0).astype(np.float32)


def cols(t, stride):
    """Deterministic frame subset kept in the fixture: first 3, last 3, every ``stride``-th."""
    sel = sorted(set(list(range(0, t, stride)) + [0, 1, 2, t - 3, t - 2, t - 1]))
    return np.array([i for i in sel if 0 <= i < t], dtype=np.int64)


def thin(d, stride=8, min_frames=64):
    """Replace every array whose last axis (time) is long by its ``cols`` subset; the
    selection is stored beside it as ``<name>__cols`` (fixtures stay small)."""
    out = {}
    for k, v in d.items():
        v = np.asarray(v)
        if v.ndim >= 2 and v.shape[-1] >= min_frames and not k.startswith(("wav_", "db_")):
            sel = cols(v.shape[-1], stride)
            out[k] = np.ascontiguousarray(v[..., sel])
            out[k + "__cols"] = sel
            out[k + "__shape"] = np.array(v.shape, dtype=np.int64)
        elif v.ndim == 1 and v.shape[0] > 20000 and not k.startswith("wav_"):
            sel = cols(v.shape[0], 16)
            out[k] = v[sel]
            out[k + "__cols"] = sel
            out[k + "__shape"] = np.array(v.shape, dtype=np.int64)
        else:
            out[k] = v
    return out


def features_goldens(sp, ft, x, msop):
    """The entries of ``features_msop.npz``: the reference's python wrappers (``sp`` = spectrum.py, ``ft`` = features.py)
    around the ``mindspore.dataset.audio`` ops -- restated ones here (``msop=1``), the real MindSpore 2.3.0 binary in
    ``oracle/make_goldens_with_mindspore.py`` (``msop=0``).  ``x``: BAC009S0002W0122 as ``io.read`` returns it."""
    f = {"msop": np.int32(msop)}
    f["spectrogram_default"] = sp.spectrogram(x)                                # (201, 480)
    f["spectrogram_512_mag"] = sp.spectrogram(x.astype(np.float32), n_fft=512, hop_length=128, power=1.0,
                                              normalized=True, window="hamming")
    f["melspectrogram_default"] = sp.melspectrogram(x)                         # (128, 480)
    f["melspectrogram_slaney"] = sp.melspectrogram(x, n_fft=512, n_mels=40, norm="slaney", mel_type="slaney",
                                                   f_min=50.0, f_max=7600.0)
    f["melscale_1024"] = sp.melscale(sp.spectrogram(x, n_fft=1024), n_stft=1024 // 2 + 1)
    f["fbank_cfg1"] = ft.fbank(x, n_mels=80, n_fft=400, hop_length=160)        # (80, 600)
    assert f["fbank_cfg1"].shape == (80, 600)
    xe = synth(4, (4, 48000))
    f["fbank_ecapa_syn4"] = ft.fbank(xe, deltas=False, n_mels=80, left_frames=0, right_frames=0,
                                     n_fft=400, hop_length=160)                  # ECAPA call, (4, 80, 301)
    xm = synth(11, (2, 16000))
    f["fbank_default_dc_syn11"] = ft.fbank(xm, deltas=True, context=True)
    f["mfcc_default_syn11"] = ft.mfcc(xm)                                      # (2, 660, 81)
    f["mfcc_cfg4"] = ft.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160)
    f["mfcc_cfg4_logmels"] = ft.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160,
                                     log_mels=True)
    f["deltas_syn"] = ft.compute_deltas(f["fbank_cfg1"][:, :100], win_length=7, pad_mode="reflect")
    f["context_3_5"] = ft.context_window(f["fbank_cfg1"][:10, :60].astype(np.float32), 3, 5)
    return f


def main():
    sys.path.insert(0, REPO)
    from oracle import ref_loader
    io, sp, ft = ref_loader.load_data_modules()
    conf = ref_loader.load_conformer_frontend()
    InputNormalization = ref_loader.load_input_normalization()
    load_json_cmvn = ref_loader.load_json_cmvn()
    os.makedirs(OUT, exist_ok=True)

    x, sr = io.read(ref_loader.sample_wav())            # float64 = int16 / 32768 (io.py:741-745)
    assert sr == 16000 and x.shape == (95984,)
    i16 = np.round(x * 32768).astype(np.int16)
    assert np.array_equal(i16.astype(np.float64) / 32768, x)

    # ---------------- spectrum (in-repo python: pinned) -----------------
    g = {"wav_i16": i16}
    g["stft_default"] = sp.stft(x)                                             # (257, 750) notebook cell 27
    assert g["stft_default"].shape == (257, 750)
    g["stft_512_256"] = sp.stft(x, n_fft=512, hop_length=256)
    g["stft_ds2"] = sp.stft(x, n_fft=320, hop_length=160, win_length=320)      # deepspeech2/dataset.py:36-41
    g["stft_400_reflect"] = sp.stft(x, n_fft=400, hop_length=160, pad_mode="reflect")
    g["stft_win400_hamming"] = sp.stft(x, n_fft=512, win_length=400, hop_length=160, window="hamming")
    g["stft_nocenter_ri"] = sp.stft(x[:20000], center=False, return_complex=False)
    xs = synth(2, (4, 16000))
    g["stft_batch_syn2"] = sp.stft(xs, n_fft=512, hop_length=256)              # cfg2 shape family
    rt = sp.istft(g["stft_default"])
    assert np.allclose(x[: rt.shape[0]], rt)                                    # tests/test_spectrum.py:38-41
    g["istft_default"] = rt
    g["istft_len90000"] = sp.istft(g["stft_default"], length=90000)
    g["istft_len99000"] = sp.istft(g["stft_default"], length=99000)
    g["istft_ds2"] = sp.istft(g["stft_ds2"], hop_length=160)
    mag1, ph1 = sp.magphase(g["stft_ds2"], 1.0)
    mag2, _ = sp.magphase(g["stft_ds2"], 2.0)
    g["magphase_ds2_mag1"], g["magphase_ds2_phase"], g["magphase_ds2_mag2"] = mag1, ph1, mag2
    kat = np.array([[3 + 4j, 0], [0, -5j]], dtype=np.complex64)
    km, kp = sp.magphase(kat, 2.0)
    g["magphase_kat_in"], g["magphase_kat_mag"], g["magphase_kat_phase"] = kat, km, kp
    # DeepSpeech2 chain: log1p + scalar norm (deepspeech2/dataset.py:43-47)
    m = np.log1p(mag1)
    g["ds2_norm"] = (m - m.mean()) / m.std()
    rng = np.random.default_rng(7)
    for nd, shape in (("2d", (40, 50)), ("3d", (3, 40, 50)), ("4d", (2, 3, 20, 30))):
        a = (rng.random(shape) ** 8 * np.exp(rng.normal(0, 4, shape[:-2] + (1, 1)))).astype(np.float64)
        g["db_in_" + nd] = a
        g["db_power_" + nd] = sp.amplitude_to_dB(a)
        g["db_mag_" + nd] = sp.amplitude_to_dB(a, stype="magnitude", ref=2.0, top_db=60.0)
    g["db_notop_2d"] = sp.amplitude_to_dB(g["db_in_2d"], top_db=None)
    g["db2amp"] = sp.dB_to_amplitude(g["db_power_2d"], 0.5, 0.5)
    np.savez_compressed(os.path.join(OUT, "spectrum.npz"), **thin(g))

    # ---------------- features via ms-op shim (msop=1: unpinned vs MindSpore binary) ----------
    f = features_goldens(sp, ft, x, msop=1)
    np.savez_compressed(os.path.join(OUT, "features_msop.npz"), **thin(f))

    # ---------------- conformer front-end + CMVN (in-repo python: pinned) --------------
    c = {}
    wav = x * (1 << 15)                                                          # conformer/dataset.py:389-390
    feats = conf.compute_fbank_feats(wav, 16000, 25, 10, 80)
    assert feats.shape == (598, 80)
    c["conformer_fbank"] = feats
    c["kaldi_mel_banks"] = conf.get_mel_banks(80, 512, 16000, 20, 8000)[0]
    xs3 = [np.round(synth(3, (n,)) * 32768).astype(np.float64) for n in (16000, 23456, 400, 559, 560)]
    fl = [conf.compute_fbank_feats(w, 16000, 25, 10, 80) for w in xs3]
    for i, ff in enumerate(fl):
        c["conformer_syn3_%d" % i] = ff
    # global CMVN: accumulate as compute_cmvn_stats.py:61-63,108-112; load as load_files.py:9-29
    s1, s2, n = np.zeros(80), np.zeros(80), 0
    for ff in [feats] + fl:
        s1 += np.sum(ff, axis=0)
        s2 += np.sum(np.square(ff), axis=0)
        n += ff.shape[0]
    c["cmvn_mean_stat"], c["cmvn_var_stat"], c["cmvn_frame_num"] = s1, s2, np.int64(n)
    tmp = os.path.join(OUT, "_cmvn_tmp.json")
    with open(tmp, "w") as fh:
        fh.write(json.dumps({"mean_stat": s1.tolist(), "var_stat": s2.tolist(), "frame_num": int(n)}))
    cm = load_json_cmvn(tmp)
    os.remove(tmp)
    c["cmvn_mean"], c["cmvn_istd"] = cm[0], cm[1]
    xf32 = feats.astype(np.float32)
    c["global_cmvn_applied"] = (xf32 - cm[0].astype(np.float32)) * cm[1].astype(np.float32)  # cmvn.py:33-36
    # utterance CMVN: InputNormalization (spec_augment.py:22-70), sentence level
    batch = np.stack([feats[:300], feats[298:598]]).copy()
    c["utt_cmvn_mean_only"] = InputNormalization(mean_norm=True, std_norm=False, norm_type="sentence").construct(batch.copy())
    c["utt_cmvn_mean_std"] = InputNormalization(mean_norm=True, std_norm=True, norm_type="sentence").construct(batch.copy())
    np.savez_compressed(os.path.join(OUT, "conformer_cmvn.npz"), **c)

    for name in sorted(os.listdir(OUT)):
        print(name, os.path.getsize(os.path.join(OUT, name)))


if __name__ == "__main__":
    main()
