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


def _mutate(rng, blob):
    import struct
    b = bytearray(blob)
    op = rng.integers(5)
    if op == 0 and len(b) > 12:          # flip bytes in the header area
        for _ in range(rng.integers(1, 4)):
            b[rng.integers(0, min(len(b), 80))] = rng.integers(256)
    elif op == 1:                        # truncate
        b = b[:rng.integers(0, len(b) + 1)]
    elif op == 2:                        # append junk
        b += bytes(rng.integers(0, 256, rng.integers(1, 40)).astype(np.uint8))
    elif op == 3 and len(b) > 44:        # change a size field
        pos = int(rng.choice([4, 16, 40]))
        b[pos:pos + 4] = struct.pack("<I", int(rng.integers(0, 2 * len(b))))
    return bytes(b)


def _outcome(fn):
    try:
        return ("ok", fn())
    except Exception as ex:   # noqa: BLE001 -- the class is what is compared
        return ("exc", type(ex).__name__)


@pytest.mark.parametrize("filelike", [True, False])
def test_fuzz_native_walk_vs_oracle(filelike):
    """Mutated containers (flipped header bytes, truncations, trailing junk, wrong size fields) x offset x duration:
    the native walk fails with the same exception class as the restated reader, or returns the same items from the
    same bytes.  (The restated reader was run against the reference's own ``read`` on 12 000 such mutations.)"""
    from mindaudio_b200 import _lib as L
    from mindaudio_b200.data import io as P
    rng = np.random.default_rng(77 + filelike)
    base = [v[0] for v in W.corpus().values()] + [v[0] for v in W.bad_corpus().values()]
    n_ok = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for it in range(1500):
            b = _mutate(rng, base[rng.integers(len(base))])
            off = float(rng.choice([0.0, 0.0, 0.003, 0.01]))
            dur = [None, None, 0.01, 0.05][rng.integers(4)]
            want = _outcome(lambda: R.wav_read(b, off, dur, filelike))
            got = _outcome(lambda: P.wav_info(b, off, dur, filelike))
            assert want[0] == got[0], (it, want, got)
            if want[0] == "exc":
                assert want[1] == got[1], (it, want, got)
                continue
            n_ok += 1
            audio, info = want[1][0], got[1]
            assert info.n_items == audio.size, it
            e = ">" if info.big_endian else "<"
            k = info.sample_kind
            if k == L.WAV_I16:
                raw = (audio * 32768).astype("<i2") if e == "<" else audio
            elif k == L.WAV_I32:
                raw = (audio * 2147483648).astype("<i4") if e == "<" else audio
            elif k in (L.WAV_U8, L.WAV_I8, L.WAV_F32, L.WAV_F64, L.WAV_I64):
                raw = audio
            else:
                continue              # 3/5/6/7-byte containers: rearranged bytes, covered by the corpus cases
            have = np.frombuffer(b, dtype=np.uint8, count=raw.nbytes, offset=int(info.data_offset)) if raw.nbytes else b""
            assert bytes(have) == raw.tobytes(), it
    assert n_ok > 200


def test_batch_walk_packs_payloads():
    """mafe_wav_stage (host threads): infos / payload offsets of many files, payloads packed back to back; the first
    malformed file is reported with the reference's exception class."""
    import ctypes as C
    from mindaudio_b200.data import io as P
    names = ["pcm16", "pcm24", "float32", "pcm16_chunks", "pcm8", "pcm16_empty", "pcm32_be", "pcm16_truncated"]
    blobs = [W.corpus()[n][0] for n in names] * 5
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        walk = P._BatchWalk(blobs, list(range(len(blobs)))).walk()
    assert len(w) == 10 and all(issubclass(x.category, P.WavFileWarning) for x in w)      # chunks + truncated, x5
    total = int(walk.offsets[-1])
    stage = np.full(total + 8, 0xAB, dtype=np.uint8)
    walk.pack(C.c_void_p(stage.ctypes.data), total)
    assert (stage[total:] == 0xAB).all()
    for k, (b, i) in enumerate(zip(blobs, walk.infos)):
        audio, sr, _ = R.wav_read(b)
        assert i.n_items == audio.size and i.sample_rate == sr
        nb = int(walk.offsets[k + 1] - walk.offsets[k])
        assert nb == audio.size * {"pcm24": 3}.get(names[k % len(names)], GOLD[names[k % len(names)]].dtype.itemsize if names[k % len(names)] not in ("pcm16", "pcm16_chunks", "pcm16_empty", "pcm16_truncated") else 2)
        assert bytes(stage[walk.offsets[k]: walk.offsets[k + 1]]) == b[i.data_offset: i.data_offset + nb]
    with pytest.raises(ValueError):
        walk.pack(C.c_void_p(stage.ctypes.data), total - 1)                               # staging buffer too small
    bad = list(blobs[:6])
    bad[4] = W.bad_corpus()["not_wave"][0]
    bad[5] = W.bad_corpus()["mulaw"][0]
    with pytest.raises(TypeError, match="file 4"):
        P._BatchWalk(bad, list(range(6))).walk()
    assert P._BatchWalk([], []).walk().offsets.tolist() == [0]


def test_file_walk_matches_bytes_walk(tmp_path):
    """mafe_wav_files_*: paths are mapped / walked / packed (pread) by the library's host threads -- same infos, same
    payload offsets, same packed bytes as handing over the files' contents; OS and format errors name the file."""
    import ctypes as C
    from mindaudio_b200.data import io as P
    names = ["pcm16", "pcm24", "float32", "pcm16_chunks", "pcm8", "pcm16_empty", "pcm32_be", "pcm16_odd_payload"]
    blobs = [W.corpus()[n][0] for n in names] * 3
    paths = []
    for k, b in enumerate(blobs):
        paths.append(str(tmp_path / ("%d.wav" % k)))
        with open(paths[-1], "wb") as fh:
            fh.write(b)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = P._BatchWalk(blobs, paths).walk()
        f = P._FileWalk([paths[0], tmp_path / "1.wav"] + paths[2:]).walk()        # str and PathLike
    assert np.array_equal(a.offsets, f.offsets)
    assert all(bytes(x) == bytes(y) for x, y in zip(a.infos, f.infos))
    total = int(a.offsets[-1])
    s1, s2 = np.zeros(total, np.uint8), np.full(total + 4, 0xCD, np.uint8)
    a.pack(C.c_void_p(s1.ctypes.data), total)
    f.pack(C.c_void_p(s2.ctypes.data), total)
    assert np.array_equal(s1, s2[:total]) and (s2[total:] == 0xCD).all()
    with pytest.raises(ValueError):
        f.pack(C.c_void_p(s2.ctypes.data), total - 1)
    f.close()
    f.close()                                                                      # idempotent
    with pytest.raises(OSError, match="nonexistent"):
        P._FileWalk(paths[:2] + [str(tmp_path / "nonexistent.wav")]).walk()
    bad = tmp_path / "bad.wav"
    bad.write_bytes(W.bad_corpus()["mulaw"][0])
    with pytest.raises(ValueError, match="bad.wav"):
        P._FileWalk(paths[:2] + [str(bad)]).walk()
    empty = tmp_path / "empty.wav"
    empty.write_bytes(b"")
    with pytest.raises(ValueError, match="empty.wav"):
        P._FileWalk([str(empty)]).walk()
    assert P._FileWalk([]).walk().offsets.tolist() == [0]


def test_negative_duration_reads_to_the_end():
    """the reference does not raise on duration < 0 -- ``count = int(duration * samplerate)`` goes negative and
    ``np.fromfile`` reads the whole chunk (io.py:500-503).  The native walk mirrors that."""
    from mindaudio_b200.data import io as P
    blob = W.corpus()["pcm16"][0]
    full = P.wav_info(blob, 0.0, None, False).n_items
    assert P.wav_info(blob, 0.0, -0.5, False).n_items == full == R.wav_read(blob, 0, -0.5, False)[0].size
