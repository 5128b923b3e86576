"""oracle/restated.py against the frozen outputs of the reference's own python (CPU, no GPU)."""
import numpy as np
import pytest

from oracle import restated as R


def close(a, b, tol=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size:
        assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b)))


STFT_CASES = {
    "stft_default": dict(),
    "stft_512_256": dict(n_fft=512, hop_length=256),
    "stft_ds2": dict(n_fft=320, hop_length=160, win_length=320),
    "stft_400_reflect": dict(n_fft=400, hop_length=160, pad_mode="reflect"),
    "stft_win400_hamming": dict(n_fft=512, win_length=400, hop_length=160, window="hamming"),
}


@pytest.mark.parametrize("name", sorted(STFT_CASES))
def test_stft(golden, name):
    out = R.stft(golden.wav(), **STFT_CASES[name])
    assert out.dtype == np.complex64
    close(golden.take("spectrum/" + name, out), golden["spectrum/" + name], 1e-7)


def test_stft_shapes_pinned_by_reference_notebook(golden):
    # tutorials/audio_data_processing_with_mindaudio.ipynb cells 27, 29; README.md:83
    assert R.stft(golden.wav(), n_fft=512).shape == (257, 750)
    assert R.magphase(R.stft(golden.wav(), n_fft=512), 1)[0].shape == (257, 750)


def test_stft_misc(golden):
    x = golden.wav()
    close(golden.take("spectrum/stft_nocenter_ri", R.stft(x[:20000], center=False, return_complex=False)),
          golden["spectrum/stft_nocenter_ri"], 1e-7)
    from tests.util import synth
    close(R.stft(synth(2, (4, 16000)), n_fft=512, hop_length=256), golden["spectrum/stft_batch_syn2"], 1e-7)
    with pytest.raises(ValueError):
        R.stft(np.zeros(100), n_fft=512)


def test_istft_roundtrip_reference_assertion(golden):
    # tests/test_spectrum.py:38-41
    x = golden.wav()
    res = R.istft(R.stft(x))
    assert res.shape == (95872,)
    assert np.allclose(x[: res.shape[0]], res)


def test_istft(golden):
    # The goldens were frozen under numpy 2.3, whose irfft transforms complex64 input in SINGLE precision;
    # the reference's own environment (numpy 1.x) and the oracle use float64 -> agreement to float32
    # rounding, except in the zero-padded tail of length > natural (y / w^2 with w -> 0 amplifies it).
    x = golden.wav()
    s = R.stft(x)
    close(golden.take("spectrum/istft_default", R.istft(s)), golden["spectrum/istft_default"], 2e-6)
    close(golden.take("spectrum/istft_len90000", R.istft(s, length=90000)), golden["spectrum/istft_len90000"], 2e-6)
    keep = golden["spectrum/istft_len99000__cols"] < 95872
    close(golden.take("spectrum/istft_len99000", R.istft(s, length=99000))[keep], golden["spectrum/istft_len99000"][keep], 2e-6)
    s2 = R.stft(x, n_fft=320, hop_length=160, win_length=320)
    close(golden.take("spectrum/istft_ds2", R.istft(s2, hop_length=160)), golden["spectrum/istft_ds2"], 2e-6)


def test_magphase_and_ds2_chain(golden):
    s = R.stft(golden.wav(), n_fft=320, hop_length=160, win_length=320)
    m1, p1 = R.magphase(s, 1.0)
    m2, _ = R.magphase(s, 2.0)
    close(golden.take("spectrum/magphase_ds2_mag1", m1), golden["spectrum/magphase_ds2_mag1"], 1e-7)
    close(golden.take("spectrum/magphase_ds2_mag2", m2), golden["spectrum/magphase_ds2_mag2"], 1e-7)
    close(golden.take("spectrum/magphase_ds2_phase", p1), golden["spectrum/magphase_ds2_phase"], 1e-7)
    km, kp = R.magphase(golden["spectrum/magphase_kat_in"], 2.0)
    assert np.array_equal(km, golden["spectrum/magphase_kat_mag"])
    assert np.array_equal(kp, golden["spectrum/magphase_kat_phase"])
    assert km[0, 0] == 25 and kp[0, 1] == 1 + 0j
    close(golden.take("spectrum/ds2_norm", R.scalar_norm(m1)), golden["spectrum/ds2_norm"], 1e-6)


@pytest.mark.parametrize("nd", ["2d", "3d", "4d"])
def test_amplitude_to_db(golden, nd):
    a = golden["spectrum/db_in_" + nd]
    close(R.amplitude_to_dB(a), golden["spectrum/db_power_" + nd])
    close(R.amplitude_to_dB(a, stype="magnitude", ref=2.0, top_db=60.0), golden["spectrum/db_mag_" + nd])


def test_db_misc(golden):
    a = golden["spectrum/db_in_2d"]
    close(R.amplitude_to_dB(a, top_db=None), golden["spectrum/db_notop_2d"])
    close(R.dB_to_amplitude(golden["spectrum/db_power_2d"], 0.5, 0.5), golden["spectrum/db2amp"])
    with pytest.raises(UserWarning):
        R.amplitude_to_dB(np.ones((2, 2), dtype=np.complex64))
    # 3-D batch couples utterances through the clamp floor (spectrum.py:81-86)
    b = golden["spectrum/db_in_3d"]
    assert not np.allclose(R.amplitude_to_dB(b), np.stack([R.amplitude_to_dB(m) for m in b]))


def test_features_msop(golden):
    from tests.util import synth
    x = golden.wav()
    g = lambda n: golden["features_msop/" + n]
    t = lambda n, full: golden.take("features_msop/" + n, full)
    close(t("spectrogram_default", R.spectrogram(x)), g("spectrogram_default"))
    close(t("spectrogram_512_mag", R.spectrogram(x.astype(np.float32), n_fft=512, hop_length=128, power=1.0,
                                                 normalized=True, window="hamming")), g("spectrogram_512_mag"), 1e-6)
    close(t("melspectrogram_default", R.melspectrogram(x)), g("melspectrogram_default"))
    close(t("melspectrogram_slaney", R.melspectrogram(x, n_fft=512, n_mels=40, norm="slaney", mel_type="slaney",
                                                     f_min=50.0, f_max=7600.0)), g("melspectrogram_slaney"))
    close(t("fbank_cfg1", R.fbank(x, n_mels=80, n_fft=400, hop_length=160)), g("fbank_cfg1"))
    assert R.fbank(x, n_mels=80, n_fft=400, hop_length=160).shape == (80, 600)
    close(t("fbank_ecapa_syn4", R.fbank(synth(4, (4, 48000)), n_mels=80, n_fft=400, hop_length=160,
                                        left_frames=0, right_frames=0)), g("fbank_ecapa_syn4"), 1e-6)
    xm = synth(11, (2, 16000))
    close(t("fbank_default_dc_syn11", R.fbank(xm, deltas=True, context=True)), g("fbank_default_dc_syn11"), 1e-6)
    out = R.mfcc(xm)
    assert out.shape == (2, 660, 81)          # docstring says 101 frames (features.py:326-329): stale
    close(t("mfcc_default_syn11", out), g("mfcc_default_syn11"), 1e-6)
    close(t("mfcc_cfg4", R.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160)), g("mfcc_cfg4"))
    close(t("mfcc_cfg4_logmels", R.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160,
                                        log_mels=True)), g("mfcc_cfg4_logmels"))
    with pytest.raises(ValueError):
        R.mfcc(xm, n_mels=10, n_mfcc=20)
    fb = R.fbank(x, n_mels=80, n_fft=400, hop_length=160)
    close(R.compute_deltas(fb[:, :100], win_length=7, pad_mode="reflect")[..., golden["features_msop/deltas_syn__cols"]],
          g("deltas_syn"))
    c = R.context_window(fb[:10, :60].astype(np.float32), 3, 5)
    assert c.shape == (1, 90, 60) and np.array_equal(c, g("context_3_5"))


def test_conformer_fbank_and_kats(golden):
    from tests.util import synth
    x = golden.wav() * (1 << 15)
    out = R.conformer_fbank(x)
    assert out.shape == (598, 80)
    close(out, golden["conformer_cmvn/conformer_fbank"], 1e-12)
    # SURVEY.md section 8c derived KATs (reference code executed)
    assert np.allclose(out[0, :5], [8.36820766, 8.82343, 8.07470698, 6.89027729, 5.54505741], atol=1e-7)
    assert np.allclose(out[300, [0, 39, 79]], [8.29245383, 10.76294849, 7.87175673], atol=1e-7)
    bank = R.kaldi_mel_banks()
    close(bank, golden["conformer_cmvn/kaldi_mel_banks"], 1e-14)
    assert bank.shape == (80, 257) and np.count_nonzero(bank) == 501 and np.all(bank[:, 256] == 0)
    assert np.count_nonzero(bank, axis=0).max() <= 2          # <= 2 filters per FFT bin
    for i, n in enumerate((16000, 23456, 400, 559, 560)):
        w = np.round(synth(3, (n,)) * 32768).astype(np.float64)
        close(R.conformer_fbank(w), golden["conformer_cmvn/conformer_syn3_%d" % i], 1e-12)
    assert R.conformer_fbank(np.ones(399)).shape == (0, 80)


def test_cmvn(golden):
    from tests.util import synth
    g = lambda n: golden["conformer_cmvn/" + n]
    feats = [g("conformer_fbank")] + [g("conformer_syn3_%d" % i) for i in range(5)]
    n, s1, s2 = R.cmvn_stats(feats)
    assert n == int(g("cmvn_frame_num"))
    close(s1, g("cmvn_mean_stat"), 1e-13)
    close(s2, g("cmvn_var_stat"), 1e-13)
    mean, istd = R.cmvn_from_stats(n, s1, s2)
    close(mean, g("cmvn_mean"), 1e-13)
    close(istd, g("cmvn_istd"), 1e-13)
    close(R.global_cmvn_apply(feats[0], mean, istd), g("global_cmvn_applied"), 1e-6)
    batch = np.stack([feats[0][:300], feats[0][298:598]])
    close(np.stack([R.utt_cmvn(b, True, False) for b in batch]), g("utt_cmvn_mean_only"), 1e-12)
    close(np.stack([R.utt_cmvn(b, True, True) for b in batch]), g("utt_cmvn_mean_std"), 1e-12)
    import json
    d = json.loads(R.cmvn_stats_json(n, s1, s2))
    assert set(d) == {"mean_stat", "var_stat", "frame_num"} and d["frame_num"] == n


def test_dither_is_deterministic_and_normal():
    a = R.dither_noise(100000, seed=1234, utt_id=7)
    b = R.dither_noise(100000, seed=1234, utt_id=7)
    c = R.dither_noise(100000, seed=1234, utt_id=8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1) < 0.02
    # Philox4x32-10 known-answer (Random123 kat_vectors: counter=0,key=0)
    w = R.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in w] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    w = R.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(v) for v in w] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_mindspore_golden_script_self_test(tmp_path):
    """oracle/make_goldens_with_mindspore.py (the script that closes the 'unpinned vs the MindSpore binary' gap where
    mindspore==2.3.0 exists): its --self-test path runs the same recipe on the restated shim and must reproduce the
    committed features_msop.npz exactly; without MindSpore the real path refuses with a clear message."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    from oracle import make_goldens_with_mindspore as mm
    rep = mm.main(["--self-test", "--out", str(tmp_path)])
    assert len(rep) >= 13 and all(v["pins_restatement"] and v["max_abs"] == 0.0 for v in rep.values())
    assert (tmp_path / "features_selftest.npz").is_file() and (tmp_path / "features_selftest_diff.json").is_file()
    try:
        import mindspore  # noqa: F401
        has_ms = not getattr(mindspore, "__file__", "").startswith(mm.HERE)
    except ImportError:
        has_ms = False
    if not has_ms:
        with pytest.raises(SystemExit):
            mm.main(["--out", str(tmp_path)])
