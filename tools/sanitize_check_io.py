"""compute-sanitizer driver for the WAV decode / resample kernels (wav.cu, resample.cu):
    compute-sanitizer --tool memcheck python tools/sanitize_check_io.py"""
import io
import sys

import numpy as np

sys.path.insert(0, '.')
import mindaudio_b200 as ma  # noqa: E402
from mindaudio_b200.data import io as P  # noqa: E402
from oracle import restated as R  # noqa: E402
from tests import wav_util as W  # noqa: E402
from tests.util import synth  # noqa: E402

for name, (blob, off, dur, fl) in W.corpus().items():
    if name in ("pcm16_be", "pcm24_be", "pcm32_be"):
        continue
    a, sr = P.read(io.BytesIO(blob), off, dur)
    if fl:
        r, _, _ = R.wav_read(blob, off, dur, True)
        assert np.array_equal(a, r), name
pcm = [np.round(synth(40 + i, (n,)) * 32768).astype(np.int16) for i, n in enumerate([16000, 901, 5361])]
blobs = [W.make_wav(pcm[0]), W.make_wav(pcm[1].astype(np.int64) << 8, width=3), W.make_wav(pcm[2], big_endian=True)]
pipe = ma.FbankPipeline()
pipe.features_from_wav([io.BytesIO(b) for b in blobs], speeds=[0.9, 1.0, 1.1])
pipe.features_from_wav([io.BytesIO(blobs[0])])
x = np.random.default_rng(0).standard_normal((3, 1237))
for new in (8000, 22050, 15999):
    assert np.abs(ma.resample(x, 16000, new) - R.resample(x, 16000, new)).max() < 1e-10
ma.pitch_shift(synth(3, (2, 5000)), 16000, 3)
print("sanitize io script ok")
