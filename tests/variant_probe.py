"""Helper of tests/test_gpu_features.py::test_kernel_variants_agree: one fixed workload through the public API, outputs
saved to an .npz.  Run in a subprocess because the library reads its A/B environment switches once per process."""
import sys

import numpy as np

sys.path.insert(0, ".")
import mindaudio_b200 as ma  # noqa: E402
from tests.util import synth  # noqa: E402


def main(out_path):
    rng = np.random.default_rng(77)
    lens = [int(v) for v in rng.integers(400, 60000, size=700)] + [400, 560, 399, 5520, 5521, 120000]
    waves = [np.round(synth(7000 + i, (m,)) * 32768).astype(np.float32) for i, m in enumerate(lens)]
    pipe = ma.FbankPipeline(cmvn="utt")
    feats, fo = pipe.features(waves, chunk_utts=len(waves))          # large batch: fused CMVN + fused frame sums by default
    x = synth(5, (3, 20000))
    mf = ma.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160)   # tensor-core DCT by default
    np.savez(out_path, feats=feats, fo=fo, mfcc=mf)


if __name__ == "__main__":
    main(sys.argv[1])
