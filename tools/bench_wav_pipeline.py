"""End-to-end input pipeline from WAV file contents (scope row f3): for every batch of `--batch` utterances,

    file bytes -> container walk + payload packing (library host threads, pinned staging) -> H2D (2 bytes / sample)
    -> fbank 80 + utterance CMVN (int16 input path) -> pad_sequence + mask on the device -> D2H of the padded batch

i.e. `FbankPipeline.features_from_wav`, what `examples/conformer/dataset.py:384-395, 456-491, 563-621` does per batch
with `read() * (1 << 15)` -> `compute_fbank_feats` -> collate.  Host wall clock around the synchronous calls.

    python tools/bench_wav_pipeline.py [--utts 2048] [--batch 128] [--steps 3]
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
from mindaudio_b200.data import io as P  # noqa: E402
from tests.wav_util import make_wav  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--files", action="store_true", help="write the WAVs to a temporary directory and pass PATHS "
                    "(mapped / walked / packed by the library's host threads) instead of the files' contents")
    args = ap.parse_args()
    rng = np.random.default_rng(8)
    lens = np.sort(rng.integers(16000, 320001, size=args.utts))            # length-bucketed batches, as the reference sorts
    blobs = [make_wav(np.clip(np.round(0.05 * 32768 * rng.standard_normal(n)), -32768, 32767).astype(np.int16)) for n in lens]
    hours = float(lens.sum()) / 16000.0 / 3600.0
    if args.files:
        import tempfile
        tmp = tempfile.mkdtemp(prefix="mafe_wav_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        paths = []
        for k, b in enumerate(blobs):
            paths.append(os.path.join(tmp, "%05d.wav" % k))
            with open(paths[-1], "wb") as fh:
                fh.write(b)
        blobs = paths
    pipe = ma.FbankPipeline(cmvn="utt")
    batches = [blobs[i:i + args.batch] for i in range(0, args.utts, args.batch)]

    def epoch(split=None):
        frames = 0
        for b in batches:
            if split is not None:
                t0 = time.perf_counter()
                P.load_batch(b)
                split[0] += time.perf_counter() - t0
            xs_pad, xs_len, xs_mask = pipe.features_from_wav(b)
            frames += int(xs_len.sum())
        return frames

    epoch()                                                              # warm-up: buffers, plans
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frames = epoch()
    dt = (time.perf_counter() - t0) / args.steps
    split = [0.0]
    epoch(split)
    if args.files:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps({"input": "paths" if args.files else "file contents (bytes)", "workload": "WAV contents (PCM16, 1-20 s, %d utterances in length-sorted batches of %d) -> padded "
                                  "fbank80 + utterance CMVN batches (features_from_wav)" % (args.utts, args.batch),
                      "audio_hours_per_s": hours / dt, "ms_per_batch": dt / len(batches) * 1e3, "frames": frames,
                      "load_batch_ms_per_batch": split[0] / len(batches) * 1e3,
                      "h2d_bytes": int(lens.sum()) * 2,
                      "timing": "host wall clock around synchronous calls, %d epochs" % args.steps}))


if __name__ == "__main__":
    main()
