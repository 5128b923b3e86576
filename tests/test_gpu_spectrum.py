"""GPU parity: mindaudio_b200.data.spectrum (CUDA, through the C ABI) vs the oracle and the goldens."""
import numpy as np
import pytest

from oracle import restated as R
from tests.util import TOL_STFT, mixed_err, stft_err, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ma():
    import __graft_entry__ as entry
    entry.build()
    import mindaudio_b200
    return mindaudio_b200


STFT_CASES = {
    "stft_default": dict(),
    "stft_512_256": dict(n_fft=512, hop_length=256),
    "stft_ds2": dict(n_fft=320, hop_length=160, win_length=320),
    "stft_400_reflect": dict(n_fft=400, hop_length=160, pad_mode="reflect"),
    "stft_win400_hamming": dict(n_fft=512, win_length=400, hop_length=160, window="hamming"),
}


@pytest.mark.parametrize("name", sorted(STFT_CASES))
def test_stft_golden(ma, golden, name):
    x = golden.wav()
    out = ma.stft(x, **STFT_CASES[name])
    assert out.dtype == np.complex64
    ref = golden["spectrum/" + name]
    assert stft_err(golden.take("spectrum/" + name, out), ref) <= TOL_STFT
    assert stft_err(out, R.stft(x, **STFT_CASES[name])) <= TOL_STFT


def test_stft_shapes_and_layout(ma, golden):
    x = golden.wav()
    s = ma.stft(x, n_fft=512)
    assert s.shape == (257, 750)                 # notebook cell 27 / README.md:83
    assert s.flags["F_CONTIGUOUS"]               # spectrum.py:249-252: order="F"
    m, p = ma.magphase(s, 1)
    assert m.shape == (257, 750) and p.shape == (257, 750)
    ri = ma.stft(x[:20000], center=False, return_complex=False)
    assert ri.shape == (257, 153, 2) and ri.dtype == np.float32
    assert stft_err(golden.take("spectrum/stft_nocenter_ri", ri), golden["spectrum/stft_nocenter_ri"]) <= TOL_STFT


@pytest.mark.parametrize("n_fft,hop,win,window,mode", [
    (512, 256, None, "hann", "constant"), (400, 160, None, "hann", "reflect"), (320, 160, 320, "hann", "constant"),
    (512, 128, 400, "hamming", "edge"), (1024, 300, 800, "blackman", "symmetric"), (2048, 512, 1200, "hann", "reflect"),
    (240, 80, None, "hann", "constant"), (600, 150, None, "bartlett", "constant"),     # 2^4*3*5, 2^3*3*5^2
    (98, 49, None, "hann", "constant"), (202, 50, None, "hann", "reflect"),             # 2*7^2 and 2*101: prime radix
    (512, 400, None, "hann", "constant"), (512, 500, None, "hann", "constant"),         # hop > n_fft/2 (reference bug)
    (512, 256, None, "hann", "wrap"),                                                    # host-pad fallback
    (256, 64, None, "hann", "reflect"), (4096, 1024, None, "hann", "constant"), (8192, 2048, 6000, "hamming", "constant"),
    (1024, 256, None, "hann", "constant"), (128, 32, None, "hann", "constant"),          # radix-16 passes: 16*16, 16^3, 16^3*2, 16*16*4; 128 = 4^3*2
    # front2048_kernel: window rows as run-time range (win 1000 -> rows 4..11), full window with zero / symmetric centre padding
    (2048, 400, 1000, "hamming", "edge"), (2048, 512, None, "hann", "constant"), (2048, 300, 1200, "hann", "symmetric"),
])
def test_stft_batch_vs_oracle(ma, n_fft, hop, win, window, mode):
    x = synth(2, (3, 16000))
    kw = dict(n_fft=n_fft, hop_length=hop, win_length=win, window=window, pad_mode=mode)
    out = ma.stft(x, **kw)
    ref = R.stft(x, **kw)
    assert out.shape == ref.shape
    for b in range(x.shape[0]):
        assert stft_err(out[b], ref[b]) <= TOL_STFT


def test_stft_float64_int_and_edge_lengths(ma):
    rng = np.random.default_rng(1)
    for length in (512, 513, 767, 768, 1024, 16001):
        x = rng.standard_normal(length)
        assert stft_err(ma.stft(x, n_fft=512, hop_length=256), R.stft(x, n_fft=512, hop_length=256)) <= TOL_STFT
        assert stft_err(ma.stft(x, n_fft=512, hop_length=256, center=False),
                        R.stft(x, n_fft=512, hop_length=256, center=False)) <= TOL_STFT
    xi = (rng.standard_normal(4000) * 1000).astype(np.int16)
    assert stft_err(ma.stft(xi), R.stft(xi)) <= TOL_STFT
    assert ma.stft(np.zeros((0, 2048), dtype=np.float32)).shape == (0, 257, 17)


@pytest.mark.parametrize("n_fft,hop,win,window,mode,center", [
    (320, 160, 320, "hann", "constant", True),      # deepspeech2: examples/deepspeech2/dataset.py:39-41
    (320, 80, None, "hann", "reflect", True), (320, 160, 200, "hamming", "edge", True), (320, 100, None, "hann", "constant", False),
    (400, 160, None, "hann", "reflect", True), (400, 200, 400, "hann", "constant", True), (400, 100, 256, "blackman", "symmetric", True),
    (400, 77, None, "hann", "constant", False),
])
def test_stft_320_400_kernel(ma, n_fft, hop, win, window, mode, center):
    """stftn16_kernel<20 / 25>: 20- / 25-point register DFT x 16, pad pass in the edge tiles, every pad mode, ragged
    lengths around the tile borders (a tile = 32 frames), single utterances shorter than one tile."""
    kw = dict(n_fft=n_fft, hop_length=hop, win_length=win, window=window, pad_mode=mode, center=center)
    for shape in ((3, 16000), (2, 32 * hop), (1, 32 * hop + n_fft - 1), (5, n_fft), (2, 1000)):
        x = synth(sum(shape), shape)
        out, ref = ma.stft(x, **kw), R.stft(x, **kw)
        assert out.shape == ref.shape and out.dtype == ref.dtype
        for b in range(shape[0]):
            assert stft_err(out[b], ref[b]) <= TOL_STFT, (shape, b)
    x1 = synth(9, (4001,)).astype(np.float64)
    assert stft_err(ma.stft(x1, **kw), R.stft(x1, **kw)) <= TOL_STFT
    mag, _ = ma.magphase(ma.stft(x1, **kw), 1.0)
    assert np.max(np.abs(mag - np.abs(R.stft(x1, **kw)))) <= TOL_STFT * np.max(np.abs(R.stft(x1, **kw)))


def test_istft_roundtrip_reference_assertion(ma, golden):
    x = golden.wav()
    res = ma.istft(ma.stft(x))
    assert res.shape == (95872,) and res.dtype == np.float64
    assert np.allclose(x[: res.shape[0]], res, atol=2e-8)       # tests/test_spectrum.py:38-41 (FP32 pipeline)
    assert np.max(np.abs(x[: res.shape[0]] - res)) <= 1e-6 * np.max(np.abs(x))


def test_istft_vs_oracle(ma, golden):
    x = golden.wav()
    s = R.stft(x)
    for kw in (dict(), dict(length=90000), dict(length=99000)):
        ref = R.istft(s, **kw)                       # float64 oracle; the device istft is float64 too
        got = ma.istft(s, **kw)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= 1e-9 * np.max(np.abs(ref))
    assert np.max(np.abs(golden.take("spectrum/istft_default", ma.istft(s)) - golden["spectrum/istft_default"])) <= 2e-6 * 0.05
    s2 = R.stft(x, n_fft=320, hop_length=160, win_length=320)
    ref = R.istft(s2, hop_length=160)
    assert np.max(np.abs(ma.istft(s2, hop_length=160) - ref)) <= 1e-9 * np.max(np.abs(ref))
    xb = synth(5, (3, 8000))
    sb = R.stft(xb, n_fft=400, hop_length=100, center=False)
    ref = R.istft(sb, hop_length=100, center=False)
    assert np.max(np.abs(ma.istft(sb, hop_length=100, center=False) - ref)) <= 1e-9 * np.max(np.abs(ref))


def test_magphase(ma, golden):
    s = R.stft(golden.wav(), n_fft=320, hop_length=160, win_length=320)
    for power in (1.0, 2.0, 0.5):
        m, p = ma.magphase(s, power)
        rm, rp = R.magphase(s, power)
        assert m.dtype == np.float32 and p.dtype == np.complex64
        assert np.max(np.abs(m - rm)) <= 1e-5 * np.max(np.abs(rm))
        assert np.max(np.abs(p - rp)) <= 1e-5
    km, kp = ma.magphase(golden["spectrum/magphase_kat_in"], 2.0)
    assert np.allclose(km, golden["spectrum/magphase_kat_mag"]) and np.allclose(kp, golden["spectrum/magphase_kat_phase"])
    assert km[0, 0] == 25 and kp[0, 1] == 1 + 0j                  # zeros -> phase 1+0j
    m3, p3 = ma.magphase(np.stack([s, s]), 1.0)                  # N-D superset
    assert m3.shape == (2,) + s.shape
    mr, ang = ma.magphase(R.stft(golden.wav()[:8000], return_complex=False), 1.0, iscomplex=False)
    rm, ra = R.magphase_real(R.stft(golden.wav()[:8000], return_complex=False), 1.0)
    assert np.max(np.abs(mr - rm)) <= 1e-5 * np.max(rm)
    big = rm > 1e-3 * rm.max()
    assert np.max(np.abs(ang - ra)[big]) <= 1e-4


@pytest.mark.parametrize("nd", ["2d", "3d", "4d"])
def test_amplitude_to_db(ma, golden, nd):
    a = golden["spectrum/db_in_" + nd]
    out = ma.amplitude_to_dB(a)
    assert out.dtype == np.float64 and out.shape == a.shape
    assert mixed_err(out, golden["spectrum/db_power_" + nd]) <= 1e-4
    assert mixed_err(ma.amplitude_to_dB(a, stype="magnitude", ref=2.0, top_db=60.0), golden["spectrum/db_mag_" + nd]) <= 1e-4
    assert mixed_err(ma.amplitude_to_dB(a.astype(np.float32), ref=np.max), R.amplitude_to_dB(a, ref=np.max)) <= 1e-4


def test_db_misc(ma, golden):
    a = golden["spectrum/db_in_2d"]
    assert mixed_err(ma.amplitude_to_dB(a, top_db=None), golden["spectrum/db_notop_2d"]) <= 1e-4
    d = golden["spectrum/db_power_2d"]
    ref = golden["spectrum/db2amp"]
    assert np.max(np.abs(ma.dB_to_amplitude(d, 0.5, 0.5) - ref) / np.maximum(np.abs(ref), 1e-30)) <= 1e-4
    b = golden["spectrum/db_in_3d"]
    batched = ma.amplitude_to_dB(b)
    assert not np.allclose(batched, np.stack([ma.amplitude_to_dB(m) for m in b]))    # batch coupling replicated
    assert ma.amplitude_to_dB(np.zeros((0, 4, 5))).shape == (0, 4, 5)


def test_spectrogram_melspectrogram_melscale(ma, golden):
    x = golden.wav()
    g = lambda n: golden["features_msop/" + n]
    t = lambda n, full: golden.take("features_msop/" + n, full)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    s = ma.spectrogram(x)
    assert s.shape == (201, 480) and s.dtype == np.float64
    assert rel(t("spectrogram_default", s), g("spectrogram_default")) <= 1e-5
    s = ma.spectrogram(x.astype(np.float32), n_fft=512, hop_length=128, power=1.0, normalized=True, window="hamming")
    assert s.dtype == np.float32 and rel(t("spectrogram_512_mag", s), g("spectrogram_512_mag")) <= 1e-5
    m = ma.melspectrogram(x)
    assert m.shape == (128, 480) and rel(t("melspectrogram_default", m), g("melspectrogram_default")) <= 1e-5
    m = ma.melspectrogram(x, n_fft=512, n_mels=40, norm="slaney", mel_type="slaney", f_min=50.0, f_max=7600.0)
    assert rel(t("melspectrogram_slaney", m), g("melspectrogram_slaney")) <= 1e-5
    sp = R.spectrogram(x, n_fft=1024)
    ms = ma.melscale(sp, n_stft=513)
    assert rel(t("melscale_1024", ms), g("melscale_1024")) <= 1e-5
    full = ma.spectrogram(x[:8000], onesided=False)
    assert full.shape[0] == 400 and rel(full, R.spectrogram(x[:8000], onesided=False)) <= 1e-5
    xb = synth(9, (2, 3, 4000))
    assert rel(ma.spectrogram(xb, pad=7, power=0.5), R.spectrogram(xb, pad=7, power=0.5)) <= 1e-5
    # stftn16 kernels with power output: n_fft 320 / 400, every power / normalisation / pad mode, tile borders
    for kw in (dict(n_fft=320, hop_length=160, power=1.0, normalized=True), dict(n_fft=400, hop_length=100, power=2.0, pad_mode="constant"),
               dict(n_fft=320, win_length=200, hop_length=80, power=3.0, window="hamming", pad_mode="edge"), dict(n_fft=400, center=False),
               dict(n_fft=512, hop_length=256), dict(n_fft=512, hop_length=100, power=1.0, normalized=True, pad_mode="symmetric"),
               dict(n_fft=512, win_length=400, hop_length=160, power=0.5, center=False)):   # stft512 kernel, power output
        for shape in ((3, 16000), (2, 32 * 200 + 399), (4, kw["n_fft"])):
            xs = synth(31 + shape[1] % 7, shape)
            assert rel(ma.spectrogram(xs, **kw), R.spectrogram(xs, **kw)) <= 1e-5, (kw, shape)
    # front2048_kernel: power / mel output, every power form, ragged half-tile borders (16 frames), run-time window rows
    for kw in (dict(n_fft=2048, win_length=1200, hop_length=300), dict(n_fft=2048, hop_length=512, power=1.0, pad_mode="constant"),
               dict(n_fft=2048, win_length=900, hop_length=256, power=3.0, window="hamming", pad_mode="edge")):
        for shape in ((2, 22050), (3, 15 * kw["hop_length"] + 2048), (2, 2048)):
            xs = synth(41 + shape[1] % 5, shape)
            assert rel(ma.spectrogram(xs, **kw), R.spectrogram(xs, **kw)) <= 1e-5, (kw, shape)
            for p in (2.0, 1.0):
                mk = dict(kw, power=p, n_mels=64, f_max=7000.0)
                assert rel(ma.melspectrogram(xs, **mk), R.melspectrogram(xs, **mk)) <= 1e-5, (mk, shape)


def test_phase_vocoder_and_time_stretch(ma):
    """Scope row f2: _phase_vocoder / time_stretch (augment.py:795-871), STFT -> vocoder -> ISTFT on the device.
    The reference's phase accumulator is a FLOAT32 array (np.angle of complex64) that reaches ~pi * hop * T rad: one
    float32 ulp of the phase is up to 2e-3 rad here, and a 1-ulp difference between CUDA's and numpy's atan2f can flip
    a rounding of the accumulator.  Parity is therefore judged at one phase ulp x magnitude (3e-3 of the scale); most
    elements agree to 1e-6."""
    x = synth(13, (2, 6000))
    spec = R.stft(x)
    for rate in (0.8, 1.0, 1.3, 2.0):
        out, ref = ma.augment._phase_vocoder(spec, rate), R.phase_vocoder(spec, rate)
        assert out.shape == ref.shape and out.dtype == ref.dtype
        err = np.abs(out - ref)
        assert np.max(err) <= 3e-3 * np.max(np.abs(ref)), rate
        assert np.mean(err <= 1e-5 * np.max(np.abs(ref))) >= 0.98, rate
        y, yr = ma.time_stretch(x, rate), R.time_stretch(x, rate)
        assert y.shape == yr.shape and y.dtype == yr.dtype
        assert np.max(np.abs(y - yr)) <= 3e-3 * np.max(np.abs(yr)), rate
    y1 = ma.time_stretch(x[0].astype(np.float64), 1.5)
    assert y1.shape == (4000,) and np.max(np.abs(y1 - R.time_stretch(x[0].astype(np.float64), 1.5))) <= 3e-3 * np.max(np.abs(y1))
    with pytest.raises(ValueError):
        ma.time_stretch(x, 0.0)



def test_resample_and_pitch_shift(ma):
    """Scope row f2: processing.resample (Fourier method, arbitrary lengths through Bluestein in complex128) against
    scipy.signal.resample as the reference calls it, and augment.pitch_shift = time_stretch -> resample -> crop / pad."""
    rng = np.random.default_rng(23)
    cases = [(997, 16000, 8000), (4001, 16000, 22050), (6000, 12700.3, 16000), (4096, 16000, 8000), (5000, 8000, 16000),
             (1, 16000, 48000), (3001, 44100, 16000), (250001, 16000, 15999)]
    for n, orig, new in cases:
        x = rng.standard_normal((2, n)) if n < 100000 else rng.standard_normal(n)
        out, ref = ma.resample(x, orig, new), R.resample(x, orig, new)
        assert out.shape == ref.shape and out.dtype == ref.dtype == np.float64
        assert np.max(np.abs(out - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), (n, orig, new)
    x32 = synth(29, (3, 2, 7000))
    out, ref = ma.resample(x32, 16000, 11025), R.resample(x32, 16000, 11025)      # scipy works in float32 for float32 input
    assert out.shape == ref.shape == (3, 2, 4824) and out.dtype == ref.dtype == np.float32
    assert np.max(np.abs(out - ref)) <= 1e-5 * np.max(np.abs(ref))
    assert ma.resample(x32, 16000, 16000) is x32
    with pytest.raises(NotImplementedError):
        ma.resample(x32, 16000, 8000, res_type="minddata")

    x = synth(13, (2, 6000))
    for n_steps in (4, -3, 0.5):
        y, yr = ma.pitch_shift(x, 16000, n_steps), R.pitch_shift(x, 16000, n_steps)
        assert y.shape == yr.shape and y.dtype == yr.dtype
        assert np.max(np.abs(y - yr)) <= 3e-3 * np.max(np.abs(yr)), n_steps     # the phase vocoder's tolerance (see above)
    y1 = ma.pitch_shift(x[0].astype(np.float64), 16000, 2, bins_per_octave=24)
    assert y1.shape == R.pitch_shift(x[0].astype(np.float64), 16000, 2, bins_per_octave=24).shape
