"""WAV decode ("next" row f3): the oracle's restatement of ``mindaudio/data/io.py:read`` against the reference's own
python (here) and its frozen outputs (anywhere); the native container walk (``mafe_wav_parse``, host only) against the
oracle.  The device decode is checked in ``test_gpu_wav.py``."""
import io
import os
import tempfile
import warnings

import numpy as np
import pytest

from oracle import restated as R
from tests import wav_util as W

HAVE_REF = os.path.isfile("/root/reference/mindaudio/data/io.py")
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "wav_io.npz"))
CASES = sorted(W.corpus())


def same_array(a, b):
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name", CASES)
def test_oracle_vs_golden(name):
    blob, off, dur, fl = W.corpus()[name]
    audio, sr, notes = R.wav_read(blob, off, dur, fl)
    same_array(audio, GOLD[name])
    assert sr == int(GOLD[name + "__sr"])
    assert len(notes) == int(GOLD[name + "__warn"])


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_oracle_vs_reference_read():
    from oracle.make_wav_goldens import load_reference_io, reference_read
    ref = load_reference_io()
    with tempfile.TemporaryDirectory() as tmp:
        for seed in (1, 2):
            for name, (blob, off, dur, fl) in W.corpus(seed).items():
                exp, sr, nw = reference_read(ref, blob, off, dur, fl, tmp, name)
                got, sr2, notes = R.wav_read(blob, off, dur, fl)
                same_array(got, exp)
                assert sr == sr2 and nw == len(notes), name
        for name, (blob, exc) in W.bad_corpus().items():
            with pytest.raises(exc), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref.read(io.BytesIO(blob))
            with pytest.raises(exc):
                R.wav_read(blob, filelike=True)


@pytest.mark.parametrize("name", CASES)
def test_native_walk_vs_oracle(name):
    from mindaudio_b200 import _lib as L
    from mindaudio_b200.data import io as P
    blob, off, dur, fl = W.corpus()[name]
    audio, sr, notes = R.wav_read(blob, off, dur, fl)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        info = P.wav_info(blob, off, dur, fl)
    assert info.n_items == audio.size
    assert info.sample_rate == sr
    assert info.channels == (audio.shape[1] if audio.ndim == 2 else 1)
    assert bool(info.big_endian) == blob.startswith(b"RIFX")
    flags = {"unknown": L.WAV_WARN_UNKNOWN_CHUNK, "eof": L.WAV_WARN_EOF, "incomplete": L.WAV_WARN_INCOMPLETE_ID}
    want = 0
    for n in notes:
        want |= flags[n]
    assert info.warnings == want
    assert all(issubclass(x.category, P.WavFileWarning) for x in w) and bool(w) == bool(notes)
    # the items start where the oracle's view starts: compare the first item's bytes
    if audio.size and info.sample_kind in (L.WAV_U8, L.WAV_F32, L.WAV_F64):
        first = np.frombuffer(blob, dtype=GOLD[name].dtype, count=1, offset=info.data_offset)[0]
        assert first == audio.reshape(-1)[0]


@pytest.mark.parametrize("name", sorted(W.bad_corpus()))
def test_native_walk_errors(name):
    from mindaudio_b200.data import io as P
    blob, exc = W.bad_corpus()[name]
    with pytest.raises(exc):
        R.wav_read(blob, filelike=True)
    with pytest.raises(exc):
        P.wav_info(blob, filelike=True)


@pytest.mark.parametrize("name", [n for n in CASES if GOLD[n].dtype.kind in "uf" and not n.startswith(("pcm16", "pcm24", "pcm32"))
                                  or n.endswith("_be") or n in ("pcm40", "pcm64")])
def test_read_passthrough_formats_without_device(name):
    """8-bit, IEEE float, int64 containers and RIFX integer files carry no arithmetic: ``read`` returns the file's
    bytes under the reference's dtype, which needs no GPU."""
    from mindaudio_b200.data import io as P
    blob, off, dur, fl = W.corpus()[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        audio, sr = P.read(io.BytesIO(blob), off, dur) if fl else _read_path(P, blob, off, dur)
    same_array(audio, GOLD[name])
    assert sr == int(GOLD[name + "__sr"])


def _read_path(P, blob, off, dur):
    with tempfile.NamedTemporaryFile(suffix=".wav") as fh:
        fh.write(blob)
        fh.flush()
        return P.read(fh.name, off, dur)


def test_file_object_is_rewound():
    from mindaudio_b200.data import io as P
    blob = W.corpus()["float32"][0]
    f = io.BytesIO(blob)
    P.read(f)
    assert f.tell() == 0      # io.py:738-739
