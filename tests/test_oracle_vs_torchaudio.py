"""Second opinion for the restated mindspore ops ([ms-op], parity unpinned vs the MindSpore
binary): torchaudio.functional is the lineage those ops document."""
import numpy as np
import pytest

from oracle import restated as R

torch = pytest.importorskip("torch")
F = pytest.importorskip("torchaudio.functional")


def test_spectrogram():
    from tests.util import synth
    x = synth(5, (2, 8000)).astype(np.float64)
    for n_fft, win, hop, power, norm, mode in ((400, 400, 200, 2.0, False, "reflect"), (512, 400, 160, 1.0, True, "constant"),
                                               (320, 320, 160, 2.0, False, "reflect")):
        ref = F.spectrogram(torch.from_numpy(x), pad=0, window=torch.hann_window(win, dtype=torch.float64), n_fft=n_fft,
                            hop_length=hop, win_length=win, power=power, normalized=norm, center=True, pad_mode=mode).numpy()
        got = R.spectrogram(x, n_fft, win, hop, 0, "hann", power, norm, True, mode)
        assert np.max(np.abs(got - ref)) <= 1e-10 * max(1, np.max(np.abs(ref)))


def test_melscale_fbanks_dct_deltas():
    for (n_stft, n_mels, norm, mt) in ((201, 80, None, "htk"), (257, 128, None, "htk"), (201, 40, "slaney", "slaney")):
        ref = F.melscale_fbanks(n_stft, 0.0, 8000.0, n_mels, 16000, norm, mt).numpy()
        got = R.melscale_fbanks(n_stft, 0.0, 8000.0, n_mels, 16000, norm or "none", mt)
        assert np.max(np.abs(got - ref)) < 3e-5          # torchaudio builds the table in float32
        assert np.count_nonzero(got, axis=1).max() <= 2  # <= 2 filters per FFT bin
    for norm in (None, "ortho"):
        ref = F.create_dct(20, 40, norm).numpy()
        assert np.max(np.abs(R.create_dct(20, 40, norm or "none") - ref)) < 3e-5
    x = np.random.default_rng(3).standard_normal((2, 13, 40))
    ref = F.compute_deltas(torch.from_numpy(x), win_length=5, mode="replicate").numpy()
    assert np.max(np.abs(R.compute_deltas(x, 5, "edge") - ref)) < 1e-12


def test_sliding_window_cmn_and_spectral_centroid():
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 50, 7)) * 3.0 + 1.0
    for kw in (dict(cmn_window=20, min_cmn_window=5, center=False, norm_vars=False), dict(cmn_window=15, min_cmn_window=100, center=True, norm_vars=True),
               dict(cmn_window=600, min_cmn_window=100, center=False, norm_vars=True), dict(cmn_window=8, min_cmn_window=3, center=False, norm_vars=True)):
        ref = F.sliding_window_cmn(torch.from_numpy(x), **kw).numpy()
        assert np.max(np.abs(R.sliding_window_cmn(x, **kw) - ref)) < 1e-9
    assert np.max(np.abs(R.sliding_window_cmn(x[0], 20, 5) - F.sliding_window_cmn(torch.from_numpy(x[0]), 20, 5).numpy())) < 1e-9
    from tests.util import synth
    w = synth(5, (2, 8000)).astype(np.float64)
    for n_fft, win, hop in ((400, 400, 200), (512, 400, 160)):
        ref = F.spectral_centroid(torch.from_numpy(w), 16000, pad=0, window=torch.hann_window(win, dtype=torch.float64), n_fft=n_fft,
                                  hop_length=hop, win_length=win).numpy()
        assert np.max(np.abs(R.spectral_centroid(w, 16000, n_fft, win, hop) - ref)) < 1e-7

