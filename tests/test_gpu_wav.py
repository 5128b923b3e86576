"""GPU parity of the WAV decode ("next" row f3): ``read`` against the frozen outputs of the reference's own ``read``
(bit exact: the scalings are powers of two), ``mafe_wav_decode`` through the C ABI for every container, and
WAV files -> padded conformer feature batch in one device round trip."""
import ctypes as C
import io
import tempfile
import warnings

import numpy as np
import pytest

from oracle import restated as R
from tests import wav_util as W
from tests.util import logmel_err, mixed_err, synth

pytestmark = pytest.mark.gpu

GOLD = np.load(__file__.replace("test_gpu_wav.py", "golden/wav_io.npz"))


@pytest.fixture(scope="module")
def ma():
    import __graft_entry__ as entry
    entry.build()
    import mindaudio_b200
    return mindaudio_b200


@pytest.mark.parametrize("name", sorted(W.corpus()))
def test_read_vs_reference_golden(ma, name):
    from mindaudio_b200.data import io as P
    blob, off, dur, fl = W.corpus()[name]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        if fl:
            audio, sr = P.read(io.BytesIO(blob), off, dur)
        else:
            with tempfile.NamedTemporaryFile(suffix=".wav") as fh:
                fh.write(blob)
                fh.flush()
                audio, sr = P.read(fh.name, off, dur)
    ref = GOLD[name]
    assert audio.dtype == ref.dtype and audio.shape == ref.shape and np.array_equal(audio, ref)
    assert sr == int(GOLD[name + "__sr"])
    assert bool(w) == bool(int(GOLD[name + "__warn"]))


def test_decode_kernel_all_containers(ma):
    """mafe_wav_decode through the C ABI: every container x both byte orders -> float32 / float64 / raw int16."""
    from mindaudio_b200 import _lib as L
    from mindaudio_b200._engine import get_engine
    eng = get_engine()
    r = np.random.default_rng(3)
    n = 4099
    cases = []
    for be in (False, True):
        e = ">" if be else "<"
        i16 = r.integers(-32768, 32768, n)
        i32 = r.integers(-(1 << 31), 1 << 31, n)
        f = r.standard_normal(n)
        cases += [
            (L.WAV_U8, be, r.integers(0, 256, n).astype("u1"), lambda a: a.astype(np.float64)),
            (L.WAV_I8, be, r.integers(-128, 128, n).astype("i1"), lambda a: a.astype(np.float64)),
            (L.WAV_I16, be, i16.astype(e + "i2"), lambda a: a.astype(np.float64) / 32768),
            (L.WAV_I32, be, i32.astype(e + "i4"), lambda a: a.astype(np.float64) / 2147483648),
            (L.WAV_I64, be, r.integers(-(1 << 62), 1 << 62, n).astype(e + "i8"), lambda a: a.astype(np.float64)),
            (L.WAV_F32, be, f.astype(e + "f4"), lambda a: a.astype(np.float64)),
            (L.WAV_F64, be, f.astype(e + "f8"), lambda a: a.astype(np.float64)),
        ]
        for width, kind in ((3, L.WAV_I24), (5, L.WAV_I40), (6, L.WAV_I48), (7, L.WAV_I56)):
            v = r.integers(-(1 << (8 * width - 1)), 1 << (8 * width - 1), n)
            cols = [((v >> (8 * k)) & 0xFF).astype(np.uint8) for k in range(width)]
            raw = np.stack(cols[::-1] if be else cols, axis=1).reshape(-1)
            if width == 3:
                cases.append((kind, be, raw, lambda a, v=v: (v << 8).astype(np.float64) / 2147483648))
            else:
                cases.append((kind, be, raw, lambda a, v=v, width=width: (v << (8 * (8 - width))).astype(np.float64)))
    for kind, be, arr, expect in cases:
        payload = np.frombuffer(arr.tobytes(), dtype=np.uint8)
        ref = expect(arr)
        for out_code, dt, scale in ((L.WAV_OUT_F64, np.float64, 1.0), (L.WAV_OUT_F32, np.float32, 1.0), (L.WAV_OUT_F32, np.float32, 32768.0)):
            out = np.empty(n, dtype=dt)
            with eng.lock:
                d_in, d_out = eng.buf("wave", payload.nbytes + 1), eng.buf("out", out.nbytes)
                # odd device address: the kernel assembles bytes, so payloads need no alignment
                keep = eng.h2d(C.c_void_p(d_in.value + 1), payload)
                L.check(eng.lib.mafe_wav_decode(eng.ctx, C.c_void_p(d_in.value + 1), n, kind, int(be), out_code, scale, d_out))
                eng.d2h(out, d_out)
                eng.sync()
                del keep
            assert np.array_equal(out, (ref * scale).astype(dt)), (kind, be, dt, scale)
    # raw int16 in host order from a big-endian payload
    a = r.integers(-32768, 32768, n).astype(">i2")
    payload = np.frombuffer(a.tobytes(), dtype=np.uint8)
    out = np.empty(n, dtype=np.int16)
    with eng.lock:
        d_in, d_out = eng.buf("wave", payload.nbytes), eng.buf("out", out.nbytes)
        keep = eng.h2d(d_in, payload)
        L.check(eng.lib.mafe_wav_decode(eng.ctx, d_in, n, L.WAV_I16, 1, L.WAV_OUT_I16, 1.0, d_out))
        eng.d2h(out, d_out)
        eng.sync()
    assert np.array_equal(out, a.astype(np.int16))
    with pytest.raises(ValueError):
        L.check(eng.lib.mafe_wav_decode(eng.ctx, d_in, n, L.WAV_F32, 0, L.WAV_OUT_I16, 1.0, d_out))
    with pytest.raises(ValueError):
        L.check(eng.lib.mafe_wav_decode(eng.ctx, d_in, n, 99, 0, L.WAV_OUT_F32, 1.0, d_out))


def _conformer_reference(blobs, max_len, cmvn):
    """read() * (1 << 15) -> compute_fbank_feats -> (utterance CMVN) -> pad, all through the oracle."""
    feats = []
    for b in blobs:
        audio, sr, _ = R.wav_read(b)
        f = R.conformer_fbank(np.asarray(audio, dtype=np.float64) * (1 << 15)).astype(np.float32)
        feats.append(R.utt_cmvn(f).astype(np.float32) if cmvn else f)
    return feats, R.conformer_collate_x(feats, max_len)


def test_wav_files_to_padded_batch(ma):
    """dataset.py:384-395 + 456-491 + 563-621 from file bytes: PCM16 files take the int16 path (bit-identical to
    feeding the decoded int16 arrays); mixed formats are decoded on the device."""
    lens = [16000, 5361, 400, 30000, 8000]
    pcm = [np.round(synth(90 + i, (n,)) * 32768).clip(-32768, 32767).astype(np.int16) for i, n in enumerate(lens)]
    blobs = [W.make_wav(p, extra_before=[(b"LIST", b"abc")] if i % 2 else ()) for i, p in enumerate(pcm)]
    pipe = ma.FbankPipeline(cmvn=None)
    xs_pad, xs_len, xs_mask = pipe.features_from_wav([io.BytesIO(b) for b in blobs], max_len=150)
    # same kernels, same int16 input -> identical bits to the array front door
    a_pad, a_len, a_mask = pipe.features_padded(pcm, max_len=150)
    assert np.array_equal(xs_pad, a_pad) and np.array_equal(xs_len, a_len) and np.array_equal(xs_mask, a_mask)
    feats, (r_pad, r_len, r_mask) = _conformer_reference(blobs, 150, False)
    assert np.array_equal(xs_len, r_len) and np.array_equal(xs_mask, r_mask)
    for i, f in enumerate(feats):                       # the log-mel criterion of test_gpu_features
        t = min(f.shape[0], 150)
        assert logmel_err(xs_pad[i, :t], f[:t].astype(np.float64)) <= 1.0, i
        assert not xs_pad[i, t:].any()

    # mixed containers: 24-bit, 32-bit, float32, big-endian 16-bit (unscaled by the reference, io.py:741-746) and 8-bit
    v = [p.astype(np.int64) for p in pcm]
    mixed = [W.make_wav(v[0] << 8, width=3), W.make_wav(v[1] << 16, width=4), W.make_wav(v[2] / 32768.0, width=4, tag=3),
             W.make_wav(v[3] // 4096, big_endian=True), W.make_wav(v[4] // 4096 + 8, width=1)]
    m_pad, m_len, m_mask = pipe.features_from_wav([io.BytesIO(b) for b in mixed], max_len=150)
    feats, (r_pad, r_len, r_mask) = _conformer_reference(mixed, 150, False)
    assert np.array_equal(m_len, r_len) and np.array_equal(m_mask, r_mask)
    for i, f in enumerate(feats):
        t = min(f.shape[0], 150)
        assert logmel_err(m_pad[i, :t], f[:t].astype(np.float64)) <= 1.0, i
    # the 24/32-bit/float files hold the same samples as the PCM16 ones: same features up to the f32 input rounding
    for i in range(3):
        assert np.max(np.abs(m_pad[i] - xs_pad[i])) <= 1e-4

    with pytest.raises(ValueError):
        pipe.features_from_wav([io.BytesIO(W.make_wav(pcm[0], channels=2))])
    with pytest.raises(ValueError):
        pipe.features_from_wav([io.BytesIO(W.make_wav(pcm[0], rate=8000))])
    # utterance CMVN fused, from paths on disk
    pipe_c = ma.FbankPipeline()
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i, b in enumerate(blobs):
            paths.append("%s/%d.wav" % (tmp, i))
            with open(paths[-1], "wb") as fh:
                fh.write(b)
        c_pad, c_len, _ = pipe_c.features_from_wav(paths)
    feats, (r_pad, r_len, _) = _conformer_reference(blobs, int(max(c_len)), True)
    assert np.array_equal(c_len, r_len)
    for i, f in enumerate(feats):
        if f.shape[0] > 1:                                 # one frame: std = 0, no eps in the reference (a11)
            assert mixed_err(c_pad[i, :f.shape[0]], f) <= 1e-3, i


def test_speed_perturbed_wav_batch(ma):
    """dataset.py:384-404: read() * (1 << 15) -> resample(waveform, sample_rate * speed, sample_rate) for speed != 1
    -> compute_fbank_feats, the resampling done on the device between the PCM decode and the front-end."""
    from mindaudio_b200.data import io as P
    lens, speeds = [16000, 9000, 5361, 12001], [0.9, 1.0, 1.1, 1.1]
    pcm = [np.round(synth(40 + i, (n,)) * 32768).clip(-32768, 32767).astype(np.int16) for i, n in enumerate(lens)]
    blobs = [W.make_wav(p) for p in pcm]
    feats, waves = [], []
    for b, v in zip(blobs, speeds):
        audio, sr, _ = R.wav_read(b)
        w = audio * (1 << 15)
        if v != 1.0:
            w = R.resample(w, sr * v, sr)
        waves.append(w)
        feats.append(R.conformer_fbank(w).astype(np.float32))
    wb = P.load_batch([io.BytesIO(b) for b in blobs], speeds=speeds)
    assert list(wb.lengths) == [len(w) for w in waves] and wb.dtype == 0 + __import__("mindaudio_b200")._lib.WAVE_F32
    pipe = ma.FbankPipeline(cmvn=None)
    xs_pad, xs_len, xs_mask = pipe.features_from_wav([io.BytesIO(b) for b in blobs], speeds=speeds)
    assert list(xs_len) == [f.shape[0] for f in feats]
    for i, f in enumerate(feats):
        assert logmel_err(xs_pad[i, :f.shape[0]], f.astype(np.float64)) <= 1.0, i
    # no perturbation drawn: the int16 path, identical to the un-perturbed call
    a = pipe.features_from_wav([io.BytesIO(b) for b in blobs], speeds=[1.0] * 4)
    b = pipe.features_from_wav([io.BytesIO(b) for b in blobs])
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    with pytest.raises(ValueError):
        pipe.features_from_wav([io.BytesIO(blobs[0])], speeds=[0.9, 1.1])
