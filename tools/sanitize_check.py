import numpy as np, sys
sys.path.insert(0, '.')
import mindaudio_b200 as ma
from oracle import restated as R
from tests.util import synth
rng = np.random.default_rng(1)
lens = [16000, 5361, 400, 399, 0, 561, 30000, 5120 + 400]
waves = [np.round(synth(50 + i, (n,)) * 32768).astype(np.float32) for i, n in enumerate(lens)]
for cmvn in (None, "utt"):
    pipe = ma.FbankPipeline(cmvn=cmvn)
    out, fo = pipe.features(waves, chunk_utts=3)
    out16, _ = pipe.features([w.astype(np.int16) for w in waves], chunk_utts=5)
x = synth(2, (3, 9000))
for kw in (dict(n_fft=512, hop_length=256), dict(n_fft=512, hop_length=128, pad_mode="reflect"), dict(n_fft=512, hop_length=200, center=False)):
    s = ma.stft(x, **kw); r = R.stft(x, **kw)
    assert np.abs(s - r).max() / np.abs(r).max() < 1e-5
f = ma.fbank(x[0], n_mels=80, n_fft=400, hop_length=160)
m = ma.mfcc(x)
for kw in (dict(n_fft=320, hop_length=160, win_length=320), dict(n_fft=400, hop_length=100, pad_mode="reflect"), dict(n_fft=320, hop_length=80, center=False)):
    s = ma.stft(x, **kw); r = R.stft(x, **kw)
    assert np.abs(s - r).max() / np.abs(r).max() < 1e-5
import random
pipe = ma.FbankPipeline(cmvn=None)
xs_pad, xs_len, xs_mask = pipe.features_padded([w for w in waves if len(w) >= 400], max_len=120, spec_aug_conf={"num_t_mask": 2, "num_f_mask": 2, "max_t": 50, "max_f": 10}, rng=random.Random(3))
ma.sliding_window_cmn(np.asarray(xs_pad[:2]), 20, 5, norm_vars=True)
ma.spectral_centroid(x, 16000)
ma.mfcc(x, n_fft=1024, n_mels=40, n_mfcc=13, deltas=False, context=False)
spec = R.stft(x[:, :4000], n_fft=512)
ma.hpss(spec, kernel_size=(13, 7), margin=(1.0, 3.0))
ma.hpss(np.abs(spec[0]).astype(np.float32), power=np.inf, mask=True)
ma.harmonic(x[0, :6000])
ma.time_stretch(x[:, :6000], 1.3)
ma.augment._phase_vocoder(spec, 0.8)
# round 2: the 2048-point front-end (every output kind, edge tiles, odd frame counts) and the fused-CMVN path with
# several utterances per CTA
x22 = synth(9, (2, 9001))
ma.stft(x22, n_fft=2048, hop_length=300, win_length=1200)
ma.spectrogram(x22, n_fft=2048, hop_length=512)
ma.melspectrogram(x22, n_fft=2048, win_length=1200, hop_length=300, n_mels=128, sample_rate=22050)
ma.mfcc(x22, n_fft=2048, n_mels=128, n_mfcc=64, hop_length=256, deltas=False, context=False)
many = [np.round(synth(200 + i, (int(n),)) * 32768).astype(np.float32) for i, n in enumerate(rng.integers(400, 9000, size=96))]
pipe = ma.FbankPipeline(cmvn="utt")
pipe.features(many, chunk_utts=40)
# the MFCC projection on the tensor cores (dct_mma_kernel): 80 -> 40 and a K / 16 = 3 instantiation with 13 coefficients
ma.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160)
ma.mfcc(x[0, :5000], deltas=False, context=False, n_mels=48, n_mfcc=13, hop_length=160)
# the deepspeech2 output kind (log1p|X| + scalar normalisation, moments accumulated by the transform)
ma.ds2_features([x22[0], x22[1, :4000]])
# frame-mean sums inside the persistent kernel (needs >= 12 tiles per SM): ragged, with one- and two-frame utterances
big_l = [int(v) for v in rng.integers(20000, 40000, size=420)] + [400, 560, 399, 0, 5520, 5521]
big = [np.round(synth(900 + i, (n,)) * 32768).astype(np.float32) for i, n in enumerate(big_l)]
pipe.features(big, chunk_utts=len(big))
pipe.features([w.astype(np.int16) for w in big], chunk_utts=len(big))
print("sanitize script ok")
