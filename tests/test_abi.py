"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/mafe.h
declares, and the python mirror keeps the reference's signatures / error behaviour BEFORE any
device call (no compute calls here: there is no GPU in the dev container)."""
import inspect

import numpy as np
import pytest

import __graft_entry__ as entry


@pytest.fixture(scope="module", autouse=True)
def built():
    entry.build()


def test_library_exports_every_declared_symbol():
    from mindaudio_b200 import _lib
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib._PROTOS, "no ctypes prototype for " + n
    assert lib.mafe_version() == 102


def test_import_does_not_touch_cuda_and_fails_loudly_without_gpu():
    import mindaudio_b200 as ma
    from mindaudio_b200 import _engine
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    assert _engine._engine is None                       # nothing created at import (fork safety)
    with pytest.raises((ma.MafeError, ValueError)):      # no CPU fallback: loud failure
        ma.stft(np.zeros(2048, dtype=np.float32))


def test_signatures_match_reference():
    import mindaudio_b200 as ma
    sig = lambda f: [(p.name, p.default) for p in inspect.signature(f).parameters.values()]
    E = inspect.Parameter.empty
    assert sig(ma.stft) == [("waveforms", E), ("n_fft", 512), ("win_length", None), ("hop_length", None),
                            ("window", "hann"), ("center", True), ("pad_mode", "constant"), ("return_complex", True)]
    assert sig(ma.istft) == [("stft_matrix", E), ("n_fft", None), ("win_length", None), ("hop_length", None),
                             ("window", "hann"), ("center", True), ("length", None)]
    assert sig(ma.magphase) == [("waveform", E), ("power", E), ("iscomplex", True)]
    assert sig(ma.amplitude_to_dB) == [("wavform", E), ("stype", "power"), ("ref", 1.0), ("amin", 1e-10), ("top_db", 80.0)]
    assert sig(ma.dB_to_amplitude) == [("wavform", E), ("ref", E), ("power", E)]
    assert [n for n, _ in sig(ma.spectrogram)] == ["waveforms", "n_fft", "win_length", "hop_length", "pad", "window",
                                                   "power", "normalized", "center", "pad_mode", "onesided"]
    assert dict(sig(ma.spectrogram))["pad_mode"] == "reflect" and dict(sig(ma.spectrogram))["n_fft"] == 400
    assert [n for n, _ in sig(ma.melspectrogram)][11:] == ["n_mels", "sample_rate", "f_min", "f_max", "norm", "mel_type"]
    assert sig(ma.fbank) == [("waveforms", E), ("deltas", False), ("context", False), ("n_mels", 40), ("n_fft", 400),
                             ("sample_rate", 16000), ("f_min", 0.0), ("f_max", None), ("left_frames", 5),
                             ("right_frames", 5), ("win_length", None), ("hop_length", None), ("window", "hann")]
    assert ma.fbanks is ma.fbank                          # README.md:41 spells it fbanks
    assert sig(ma.mfcc)[:5] == [("waveforms", E), ("deltas", True), ("context", True), ("n_mels", 23), ("n_mfcc", 20)]
    assert sig(ma.compute_deltas) == [("specgram", E), ("win_length", 5), ("pad_mode", "edge")]
    assert sig(ma.context_window) == [("waveforms", E), ("left_frames", 0), ("right_frames", 0)]
    import mindaudio
    assert mindaudio.stft is ma.stft and mindaudio.data.spectrum.magphase is ma.magphase


def test_python_level_errors_raise_before_any_device_call():
    import mindaudio_b200 as ma
    x = np.zeros(100, dtype=np.float32)
    with pytest.raises(ValueError):
        ma.stft(x, n_fft=512)                              # spectrum.py:182-187
    with pytest.raises(ValueError):
        ma.stft(x, n_fft=512, center=False)                # spectrum.py:243-246
    with pytest.raises(ValueError):
        ma.stft(np.zeros(4096), hop_length=0)              # spectrum.py:295-296
    with pytest.raises(ValueError):
        ma.stft(np.zeros(4096), n_fft=256, win_length=512)  # spectrum.py:331-334
    with pytest.raises(ValueError):
        ma.mfcc(np.zeros(16000), n_mels=10, n_mfcc=20)     # features.py:333-336
    with pytest.raises(UserWarning):
        ma.amplitude_to_dB(np.ones((2, 3), dtype=np.complex64))   # spectrum.py:59-64
    with pytest.raises(TypeError):
        ma.context_window(np.zeros(5))                     # features.py:113-115
    with pytest.raises(ValueError):
        ma.spectrogram(np.zeros(4096), window="nope")      # WindowType(window) coercion
    with pytest.raises(ValueError):
        ma.melspectrogram(np.zeros(4096), pad_mode="wrap")  # BorderType(pad_mode) coercion
    assert ma.WindowType("hann") is ma.WindowType.HANN and ma.BorderType.REFLECT == "reflect"


def test_host_tables_match_oracle_tables():
    from mindaudio_b200 import _tables as T
    from oracle import restated as R
    assert np.max(np.abs(T.povey_window(400) - R.povey_window(400))) < 1e-14
    assert np.max(np.abs(T.kaldi_triangle_bank(80, 512, 16000, 20, 8000) - R.kaldi_mel_banks())) < 1e-12
    a = T.hz_triangle_bank(201, 80, 16000, 0.0, 8000.0)
    assert np.max(np.abs(a - R.melscale_fbanks(201, 0.0, 8000.0, 80, 16000).T)) < 1e-12
    a = T.hz_triangle_bank(257, 40, 16000, 50.0, 7600.0, "slaney", "slaney")
    assert np.max(np.abs(a - R.melscale_fbanks(257, 50.0, 7600.0, 40, 16000, "slaney", "slaney").T)) < 1e-12
    assert np.max(np.abs(T.dct_matrix(20, 40, "ortho") - R.create_dct(20, 40, "ortho"))) < 1e-7
    assert np.max(np.abs(T.analysis_window("hamming", 400, 512) - R.pad_center(R.periodic_window("hamming", 400), 512))) == 0
