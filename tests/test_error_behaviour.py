"""Argument errors of the path: the product raises what the reference raises (class and message), before any device
work -- so this runs without a GPU; with the reference tree mounted both are called side by side."""
import os
import warnings

import numpy as np
import pytest

HAVE_REF = os.path.isdir("/root/reference/mindaudio")

X = np.random.default_rng(0).standard_normal(4000)
Z = (np.random.default_rng(1).standard_normal((257, 20)) + 1j).astype(np.complex64)

# name -> (call on a spectrum-like module, exception class, fragment of the message)
SPECTRUM_CASES = {
    "stft_short_uncentred": (lambda m: m.stft(X[:100], n_fft=512, center=False), ValueError, "too large for uncentered analysis"),
    "stft_short_reflect": (lambda m: m.stft(X[:100], n_fft=512, pad_mode="reflect"), ValueError, "too small for input signal"),
    "stft_window_longer_than_fft": (lambda m: m.stft(X, n_fft=256, win_length=400), ValueError, "Target size (256) must be at least input size (400)"),
    "db_complex_input": (lambda m: m.amplitude_to_dB(Z), UserWarning, "amplitude_to_db was called on complex input"),
    "frame_hop0": (lambda m: m.frame(X, frame_length=64, hop_length=0), ValueError, "Invalid hop_length: 0"),
    "pad_center_small": (lambda m: m._pad_center(np.ones(10), 5), ValueError, "Target size (5) must be at least input size (10)"),
    "compute_amplitude_type": (lambda m: m.compute_amplitude(X[None], amp_type="rms"), TypeError, "Unsupported amplitude type 'rms'"),
}


@pytest.mark.parametrize("name", sorted(SPECTRUM_CASES))
def test_spectrum_argument_errors(name):
    from mindaudio_b200.data import spectrum as ours
    fn, exc, frag = SPECTRUM_CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(exc) as e1:
            fn(ours)
    assert frag in str(e1.value)
    if HAVE_REF:
        from oracle import ref_loader
        _, sp, _ = ref_loader.load_data_modules()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with pytest.raises(exc) as e2:
                fn(sp)
        assert type(e1.value) is type(e2.value)
        assert frag in str(e2.value)


def test_feature_argument_errors():
    import mindaudio_b200 as ma
    with pytest.raises(ValueError, match="no more than # mel bins"):
        ma.mfcc(X[None].astype(np.float32), n_mels=20, n_mfcc=40)
    with pytest.raises(ValueError, match="no less than 3"):
        ma.compute_deltas(np.zeros((2, 40, 30), dtype=np.float32), win_length=2)
    with pytest.raises(TypeError, match="non-negative"):
        ma.soft_mask(-np.ones((4, 4)), np.ones((4, 4)))
    with pytest.raises(TypeError, match="shape mismatch"):
        ma.soft_mask(np.ones((4, 4)), np.ones((4, 5)))
    with pytest.raises(TypeError, match="strictly positive"):
        ma.soft_mask(np.ones((4, 4)), np.ones((4, 4)), power=0)
    with pytest.raises(TypeError, match="Margins must be >= 1.0"):
        ma.hpss(np.ones((257, 20), dtype=np.float32), margin=0.5)
    with pytest.raises(ValueError, match="rate must be a positive number"):
        ma.time_stretch(X, 0.0)
    with pytest.raises(ValueError, match="mask_param should be in"):
        ma.frequencymasking(np.zeros((1, 40, 30), dtype=np.float32), frequency_mask_param=100)
    with pytest.raises(ValueError, match="f_min"):
        ma.melscale(np.ones((201, 10), dtype=np.float32), f_min=9000, f_max=8000)
    with pytest.raises(RuntimeError, match="freq=201"):
        ma.melscale(np.ones((100, 10), dtype=np.float32))
    with pytest.raises(ValueError, match="Invalid hop_length"):
        ma.stft(X, hop_length=0)
    with pytest.raises(ValueError, match="expected 1 \\+ n_fft//2"):
        ma.istft(Z, n_fft=400)
    with pytest.raises(RuntimeError, match="expected at least 2 dimensions"):
        ma.sliding_window_cmn(np.zeros(10))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_feature_errors_match_reference_python():
    """soft_mask / hpss argument checks are plain python in the reference (features.py:438-528): same class, same text."""
    from oracle import ref_loader
    import mindaudio_b200 as ma
    _, _, ft = ref_loader.load_data_modules()
    calls = [lambda m: m.soft_mask(-np.ones((4, 4)), np.ones((4, 4))),
             lambda m: m.soft_mask(np.ones((4, 4)), np.ones((4, 5))),
             lambda m: m.soft_mask(np.ones((4, 4)), np.ones((4, 4)), power=0),
             lambda m: m.hpss(np.ones((257, 20), dtype=np.float32), margin=0.5)]
    for fn in calls:
        with pytest.raises(Exception) as e1:
            fn(ft)
        with pytest.raises(Exception) as e2:
            fn(ma)
        assert type(e1.value) is type(e2.value) and str(e1.value) == str(e2.value)


def test_extension_argument_errors_without_a_gpu():
    """mel_and_mfcc / mel_and_fbank (extensions over features.mfcc / fbank) refuse what mfcc / melspectrogram refuse, with the
    reference's messages (features.py:337 ff., spectrum.py:594 ff.), before any device work."""
    import mindaudio_b200 as ma
    with pytest.raises(ValueError, match="number of MFCC coefficients must be no more than"):
        ma.mel_and_mfcc(X, n_mels=20, n_mfcc=21)
    with pytest.raises(ValueError, match="should be no more than f_max"):
        ma.mel_and_fbank(X, f_min=5000.0, f_max=4000.0)
