"""Freeze what the reference's own ``mindaudio.data.io.read`` returns for the WAV corpus of ``tests/wav_util.py`` --
TEST INFRASTRUCTURE.

    python -m oracle.make_wav_goldens        # dev container only (/root/reference must exist)

Writes ``tests/golden/wav_io.npz``: ``<case>`` = the returned array (dtype and byte order kept), ``<case>__sr`` = the
returned sample rate, ``<case>__warn`` = number of WavFileWarning raised.  ``io.py`` is pure numpy, so the reference
module is executed as it is (file path input for the descriptor branch, ``io.BytesIO`` for the file-like branch).
"""
import importlib.util
import io
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_IO = "/root/reference/mindaudio/data/io.py"


def load_reference_io():
    spec = importlib.util.spec_from_file_location("_ref_wav_io", REF_IO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_read(ref, blob, offset, duration, filelike, tmpdir, name="x"):
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        if filelike:
            audio, sr = ref.read(io.BytesIO(blob), offset, duration)
        else:
            path = os.path.join(tmpdir, name + ".wav")
            with open(path, "wb") as fh:
                fh.write(blob)
            audio, sr = ref.read(path, offset, duration)
    return audio, sr, len(w)


def main():
    sys.path.insert(0, REPO)
    from tests import wav_util
    ref = load_reference_io()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (blob, off, dur, fl) in wav_util.corpus().items():
            audio, sr, nw = reference_read(ref, blob, off, dur, fl, tmp, name)
            out[name], out[name + "__sr"], out[name + "__warn"] = audio, np.int64(sr), np.int64(nw)
    path = os.path.join(REPO, "tests", "golden", "wav_io.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(wav_util.corpus()), "cases")


if __name__ == "__main__":
    main()
