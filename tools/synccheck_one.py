import numpy as np, sys
sys.path.insert(0, '.')
import mindaudio_b200 as ma
from tests.util import synth
which = sys.argv[1]
x = synth(2, (3, 9000))
if which == "h": ma.harmonic(x[0, :6000])
if which == "s2048": ma.stft(x, n_fft=2048, hop_length=300, win_length=1200)
if which == "mfcc": ma.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160)
if which == "stft512": ma.stft(x, n_fft=512, hop_length=256)
if which == "stft320": ma.stft(x, n_fft=320, hop_length=160, win_length=320)
if which == "fbank": 
    pipe = ma.FbankPipeline(cmvn="utt"); pipe.features([np.round(x[0]*32768).astype(np.float32), np.round(x[1,:5000]*32768).astype(np.float32)])
print("ok", which)
