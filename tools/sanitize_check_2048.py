"""compute-sanitizer workload for front2048_kernel alone (run under memcheck / racecheck): every output kind, the three
window-row instantiations (win 1200 -> rows 3..12, full window, run-time range), centre-padded first / last half-tiles
staged with cp.async, ragged frame counts (odd pair at the end of a half-tile)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import mindaudio_b200 as ma
from oracle import restated as R
from tests.util import synth

x = synth(9, (3, 9001))
for kw in (dict(n_fft=2048, hop_length=300, win_length=1200), dict(n_fft=2048, hop_length=512), dict(n_fft=2048, hop_length=400, win_length=1000, window="hamming", pad_mode="edge"),
           dict(n_fft=2048, hop_length=300, win_length=1200, pad_mode="constant"), dict(n_fft=2048, hop_length=256, center=False)):
    s = ma.stft(x, **kw)
    r = R.stft(x, **kw)
    assert np.abs(s - r).max() / np.abs(r).max() < 1e-5, kw
for p in (2.0, 1.0, 3.0):
    ma.spectrogram(x, n_fft=2048, hop_length=512, power=p)
    ma.melspectrogram(x, n_fft=2048, win_length=1200, hop_length=300, n_mels=128, sample_rate=22050, power=p)
ma.melspectrogram(x[0, :2048], n_fft=2048, win_length=900, hop_length=256, n_mels=40)
ma.mfcc(x, n_fft=2048, n_mels=128, n_mfcc=64, hop_length=256, deltas=False, context=False)
ma.fbank(x, n_fft=2048, n_mels=64, hop_length=512)
ma.ds2_features([x[0], x[1, :4000]], n_fft=2048, hop_length=300, win_length=1200)
ma.harmonic(x[0, :6000])
print("sanitize 2048 ok")
