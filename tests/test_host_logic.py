"""Host-side logic that needs no GPU: SpecAugment rectangle drawing (same `random` sequence as the reference), the
host route of pad_sequence (label sequences), make_pad_mask, the phase-vocoder step tables, enum coercion."""
import random

import numpy as np
import pytest

from oracle import restated as R


def test_spec_aug_rects_reproduce_reference_draws():
    from mindaudio_b200.data.masking import spec_aug_rects
    conf = {"num_t_mask": 3, "num_f_mask": 2, "max_t": 40, "max_f": 12}
    rng = np.random.default_rng(2)
    xs = [rng.standard_normal((n, 80)).astype(np.float32) + 3.0 for n in (90, 17, 333, 1)]
    ref = R.spec_aug([x.copy() for x in xs], conf, random.Random(99))
    rects = spec_aug_rects([x.shape for x in xs], conf, random.Random(99))
    assert rects.dtype == np.int32 and rects.shape[1] == 5
    out = [x.copy() for x in xs]
    for item, r0, r1, c0, c1 in rects:
        out[item][r0:r1, c0:c1] = 0
    assert all(np.array_equal(a, b) for a, b in zip(out, ref))
    assert spec_aug_rects([(10, 80)], {}, random.Random(1)).shape == (0, 5)


def test_pad_sequence_host_route_and_masks():
    from mindaudio_b200.data.collate import make_pad_mask, pad_sequence
    rng = np.random.default_rng(4)
    labels = [rng.integers(0, 50, size=n).astype(np.int32) for n in (4, 9, 2)]
    for kw in (dict(padding_value=-1, padding_max_len=10), dict(padding_value=0), dict(batch_first=False, padding_value=7, padding_max_len=3)):
        a, b = pad_sequence(labels, **kw), R.pad_sequence(labels, **kw)
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert np.array_equal(make_pad_mask([5, 3, 2]), R.make_pad_mask([5, 3, 2]))
    assert np.array_equal(make_pad_mask(np.array([5, 3, 2]), 8), R.make_pad_mask(np.array([5, 3, 2]), 8))


def test_vocoder_tables_match_numpy_arange():
    from mindaudio_b200.data.augment import _vocoder_tables
    for n_frames, rate in ((47, 0.8), (47, 1.0), (47, 1.3), (100, 2.0), (3, 7.5)):
        n_steps, phi = _vocoder_tables(n_frames, 257, rate, 128)
        assert n_steps == len(np.arange(0, n_frames, rate, dtype=np.float64))
        assert phi.shape == (257,) and phi[0] == 0.0 and np.isclose(phi[-1], np.pi * 128)


def test_enums_accept_strings_and_members():
    import mindaudio_b200 as ma
    assert ma.WindowType("hann") is ma.WindowType.HANN and ma.BorderType("reflect") is ma.BorderType.REFLECT
    with pytest.raises(ValueError):
        ma.WindowType("nope")


def test_resample_host_side_contract():
    """processing.resample / io.load_batch: everything decided on the host (lengths, argument errors) without a GPU."""
    import io as _io
    import mindaudio_b200 as ma
    from mindaudio_b200.data import io as P
    from oracle import restated as R
    from tests import wav_util as W
    rng = np.random.default_rng(2)
    for n in [1, 2, 9000, 9001, 16000, 31999, 44100] + [int(v) for v in rng.integers(1, 400000, 40)]:
        for speed in (0.9, 1.1):
            x = np.zeros(n)
            assert P.resampled_length(n, 16000 * speed, 16000) == R.resample(x, 16000 * speed, 16000).shape[-1], (n, speed)
        assert P.resampled_length(n, 44100, 16000) == R.resample(np.zeros(n), 44100, 16000).shape[-1]
    x = np.zeros((2, 100), dtype=np.float32)
    assert ma.resample(x, 8000, 8000) is x                                       # processing.py:167-168
    with pytest.raises(NotImplementedError):
        ma.resample(x, 16000, 8000, res_type="minddata")
    with pytest.raises(NotImplementedError):
        ma.resample(x.astype(np.complex64), 16000, 8000)
    stereo = W.make_wav(np.arange(20), channels=2)
    mono = W.make_wav(np.arange(20))
    with pytest.raises(ValueError, match="channels"):
        P.load_batch([_io.BytesIO(stereo)])
    with pytest.raises(ValueError, match="speed factors"):
        P.load_batch([mono], speeds=[0.9, 1.1])
    with pytest.raises(ValueError, match="positive"):
        P.load_batch([mono], speeds=[0.0])
    with pytest.raises(ValueError, match="64-bit"):
        P.load_batch([W.make_wav(np.arange(8), width=8)])
    with pytest.raises(UnboundLocalError):      # a container without chunks: what the reference's read() raises (io.py:741)
        P.load_batch([b"RIFF\x00\x00\x00\x00WAVE"])
    with pytest.raises(ValueError, match="file"):
        P.load_batch([mono, b"JUNKJUNKJUNK"])


def test_speed_perturb_draws_like_the_reference(monkeypatch):
    """augment.speed_perturb: same np.random call sequence and target frequency as augment.py:629-637 (the resampling
    itself is the device op checked in test_gpu_spectrum)."""
    import os
    from mindaudio_b200.data import augment as A
    calls = []
    monkeypatch.setattr(A, "resample", lambda w, o, n: calls.append((o, n)) or w)
    x = np.zeros((2, 1000), dtype=np.float32)
    np.random.seed(11)
    out = A.speed_perturb(x, 16000, perturb_prob=0.0)                 # rand(1) > 0 -> untouched copy
    assert out is not x and np.array_equal(out, x) and not calls
    expect = []
    np.random.seed(12)
    for _ in range(6):
        A.speed_perturb(x, 16000, speeds=[90, 100, 110])
    np.random.seed(12)
    for _ in range(6):
        assert not np.random.rand(1) > 1.0
        expect.append((16000, 16000 * [90, 100, 110][np.random.randint(0, 3, (1,))[0]] // 100))
    assert calls == expect and len({n for _, n in calls}) > 1
    if os.path.isfile("/root/reference/mindaudio/data/augment.py"):
        from oracle import ref_loader
        g = ref_loader._extract_functions("/root/reference/mindaudio/data/augment.py", ["speed_perturb"],
                                          {"resample": lambda w, o, n: ref_calls.append((o, n)) or w})
        ref_calls = []
        np.random.seed(12)
        for _ in range(6):
            g["speed_perturb"](x, 16000, speeds=[90, 100, 110])
        assert ref_calls == expect
