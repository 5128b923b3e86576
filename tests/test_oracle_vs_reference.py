"""oracle/restated.py against the reference's own python executed live (dev container only:
skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle import ref_loader, restated as R

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    io, sp, ft = ref_loader.load_data_modules()
    return dict(io=io, sp=sp, ft=ft, conf=ref_loader.load_conformer_frontend())


def test_all_sample_wavs(ref):
    import os
    d = os.path.dirname(ref_loader.sample_wav())
    for name in sorted(os.listdir(d)):
        x, sr = ref["io"].read(os.path.join(d, name))
        assert sr == 16000
        for kw in (dict(), dict(n_fft=400, hop_length=160, pad_mode="reflect"), dict(n_fft=320, hop_length=160)):
            assert np.array_equal(ref["sp"].stft(x, **kw), R.stft(x, **kw))
        a = ref["conf"].compute_fbank_feats(x * (1 << 15), 16000, 25, 10, 80)
        assert np.max(np.abs(a - R.conformer_fbank(x * (1 << 15)))) < 1e-11
        assert np.max(np.abs(ref["ft"].fbank(x, n_mels=80, n_fft=400, hop_length=160)
                             - R.fbank(x, n_mels=80, n_fft=400, hop_length=160))) < 1e-11


def test_stft_latent_bug_is_fixed_not_replicated(ref):
    # spectrum.py:237 raises AttributeError for some lengths with hop > n_fft/2
    x = np.random.default_rng(0).standard_normal(1100)
    bad = 0
    for hop in (300, 400, 500):
        try:
            a = ref["sp"].stft(x, n_fft=512, hop_length=hop)
            assert np.array_equal(a, R.stft(x, n_fft=512, hop_length=hop))
        except AttributeError:
            bad += 1
            assert R.stft(x, n_fft=512, hop_length=hop).shape[-1] == 1 + 1100 // hop
    assert bad >= 1


def test_context_window_matches_grouped_conv(ref):
    rng = np.random.default_rng(1)
    z = rng.standard_normal((3, 7, 50)).astype(np.float32)
    for l, r in ((3, 5), (4, 4), (5, 3), (0, 0), (5, 5), (0, 3), (2, 0)):
        assert np.array_equal(ref["ft"].context_window(z, l, r), R.context_window(z, l, r))
    z4 = rng.standard_normal((2, 3, 7, 20)).astype(np.float32)
    assert np.array_equal(ref["ft"].context_window(z4, 2, 3), R.context_window(z4, 2, 3))


def test_collate_matches_reference():
    """pad_sequence / make_pad_mask restatements vs the reference's own functions (bit exact)."""
    ref = ref_loader.load_collate()
    rng = np.random.default_rng(7)
    seqs = [rng.standard_normal((n, 5)).astype(np.float32) for n in (7, 3, 11, 1)]
    labels = [rng.integers(0, 50, size=n).astype(np.int32) for n in (4, 9, 2)]
    for kw in (dict(), dict(padding_max_len=9), dict(batch_first=False, padding_max_len=12), dict(padding_max_len=2)):
        a = ref.pad_sequence(seqs, padding_value=0.0, atype=np.float32, **kw)
        b = R.pad_sequence(seqs, padding_value=0.0, atype=np.float32, **kw)
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    a, b = ref.pad_sequence(labels, padding_value=-1, padding_max_len=10), R.pad_sequence(labels, padding_value=-1, padding_max_len=10)
    assert a.dtype == b.dtype and np.array_equal(a, b)
    lens = np.array([5, 3, 2], dtype=np.int32)
    assert np.array_equal(ref.make_pad_mask(lens), R.make_pad_mask(lens))
    assert np.array_equal(ref.make_pad_mask(lens, max_len=8), R.make_pad_mask(lens, max_len=8))


def test_spec_aug_matches_reference():
    """The restated SpecAugment draws the same ``random`` sequence as the reference method (bit exact)."""
    import random
    ref_fn = ref_loader.load_spec_aug()
    conf = {"num_t_mask": 2, "num_f_mask": 2, "max_t": 50, "max_f": 10}
    rng = np.random.default_rng(3)
    xs = [rng.standard_normal((n, 80)).astype(np.float32) + 5.0 for n in (120, 33, 400)]
    random.seed(1234)
    a = ref_fn(None, [x.copy() for x in xs], conf)
    b = R.spec_aug([x.copy() for x in xs], conf, random.Random(1234))
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    assert any((u == 0).any() for u in a)


def test_phase_vocoder_matches_reference(ref):
    """Restated phase vocoder == the reference's ``_phase_vocoder`` (same numpy operations: bit exact), and
    time_stretch composed from the reference's own stft / istft agrees with the restated chain."""
    pv = ref_loader.load_phase_vocoder()
    from tests.util import synth
    x = synth(13, (2, 6000)).astype(np.float64)
    spec = ref["sp"].stft(x)
    for rate in (0.8, 1.0, 1.3, 2.0):
        a, b = pv(spec, rate), R.phase_vocoder(R.stft(x), rate)
        assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)
        n = int(round(x.shape[-1] / rate))
        ya = ref["sp"].istft(a, length=n)
        yb = R.time_stretch(x, rate)
        assert ya.shape == yb.shape and np.max(np.abs(ya - yb)) <= 1e-6 * max(1.0, np.max(np.abs(ya)))


def test_hpss_matches_reference(ref):
    """Restated soft_mask / hpss / harmonic == the reference's own functions (scipy's median_filter included)."""
    from tests.util import synth
    ft = ref["ft"]
    x = synth(17, (2, 9000)).astype(np.float64)
    spec = ref["sp"].stft(x, n_fft=512)
    for kw in (dict(), dict(kernel_size=(13, 7), margin=(1.0, 3.0)), dict(power=1.0, kernel_size=8), dict(mask=True, margin=2.0)):
        a, b = ft.hpss(spec[0], **kw), R.hpss(spec[0], **kw)
        for u, v in zip(a, b):
            assert u.shape == v.shape and u.dtype == v.dtype and np.array_equal(u, v)
    mag = np.abs(spec[1]).astype(np.float32)
    a, b = ft.hpss(mag, power=np.inf, mask=True), R.hpss(mag, power=np.inf, mask=True)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    ya, yb = ft.harmonic(x[0]), R.harmonic(x[0])
    # the reference's istft runs numpy >= 2's single-precision irfft on complex64 here; the oracle transforms in float64
    assert ya.shape == yb.shape and np.max(np.abs(ya - yb)) <= 1e-6 * max(1.0, np.max(np.abs(ya)))



def test_pitch_shift_matches_reference(ref):
    """Restated resample / pitch_shift == the reference's own functions (both call scipy.signal.resample), including the
    crop / pad to the STRETCHED length (augment.py:901)."""
    ps = ref_loader.load_pitch_shift()
    from tests.util import synth
    x = synth(19, (2, 5000)).astype(np.float64)
    for orig, new in ((16000, 8000), (16000, 22050), (12700.3, 16000), (16000, 16000)):
        a, b = ps.resample(x, orig, new), R.resample(x, orig, new)
        assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)
    a32 = ps.resample(x.astype(np.float32), 16000, 11025)
    assert a32.dtype == np.float32 and np.array_equal(a32, R.resample(x.astype(np.float32), 16000, 11025))
    for n_steps in (4, -3, 0.5):
        a, b = ps.pitch_shift(x, 16000, n_steps), R.pitch_shift(x, 16000, n_steps)
        assert a.shape == b.shape and a.dtype == b.dtype
        assert np.max(np.abs(a - b)) <= 1e-6 * max(1.0, np.max(np.abs(a)))
        assert a.shape[-1] == int(round(x.shape[-1] / 2.0 ** (-n_steps / 12)))
