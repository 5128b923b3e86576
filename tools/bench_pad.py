"""Secondary measurement: the collate step behind the front-end (mafe_pad_sequence: padded batch + lengths + masks,
mindaudio/utils/common.py:10-52, examples/conformer/dataset.py:563-569) on 128 ragged utterances per batch, host call
through FbankPipeline.features_padded (front-end + CMVN + pad in one device round trip).

    python tools/bench_pad.py [--utts 128] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

import mindaudio_b200 as ma  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rng = np.random.default_rng(4)
    lens = rng.integers(16000, 320001, size=args.utts)
    waves = [np.round(np.clip(0.05 * rng.standard_normal(int(n)), -1, 1) * 32768).astype(np.float32) for n in lens]
    pipe = ma.FbankPipeline(cmvn="utt")
    for _ in range(3):
        xs_pad, xs_len, xs_mask = pipe.features_padded(waves)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        xs_pad, xs_len, xs_mask = pipe.features_padded(waves)
    dt = (time.perf_counter() - t0) / args.steps
    hours = float(lens.sum()) / 16000.0 / 3600.0
    print(json.dumps({"workload": "features_padded: fbank + utterance CMVN + pad_sequence + masks, %d ragged utterances per call (host arrays in and out)" % args.utts,
                      "ms_per_call": dt * 1e3, "audio_hours_per_s": hours / dt, "padded_shape": list(xs_pad.shape)}))


if __name__ == "__main__":
    main()
