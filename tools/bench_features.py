"""Secondary measurement: features.fbank / mfcc on the ECAPA-style configuration (BASELINE.json configs[0]/[3]:
n_fft=400, hop=160, 80 mel, hann, reflect padding, dB with the batch-wide top_db floor), device-resident.

    python tools/bench_features.py [--batch 4096] [--seconds 3] [--steps 10] [--mfcc]
"""
import argparse
import ctypes as C
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mindaudio_b200 import _lib as L, _tables as T  # noqa: E402
from mindaudio_b200._engine import get_engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--mfcc", action="store_true")
    ap.add_argument("--fastspeech2", action="store_true",
                    help="melspectrogram n_fft=2048 win=1200 hop=300 128 mel at 22 050 Hz (examples/fastspeech2): generic kernel")
    args = ap.parse_args()
    eng = get_engine()
    eng.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    n = int(args.seconds * 16000)
    bank = T.hz_triangle_bank(201, 80, 16000, 0.0, 8000.0)
    kw = dict(n_fft=400, hop=160, center=True, pad_mode="reflect", window=T.analysis_window("hann", 400, 400), power=2.0,
              mel_fb=bank, log_kind=L.LOG_DB, log_arg=1e-10, log_mult=10.0, log_offset=0.0, top_db=80.0)
    sr = 16000
    if args.fastspeech2:
        sr = 22050
        n = int(args.seconds * sr)
        bank = T.hz_triangle_bank(1025, 128, sr, 0.0, sr // 2)
        plan = eng.plan(n_fft=2048, hop=300, center=True, pad_mode="reflect", window=T.analysis_window("hann", 1200, 2048),
                        power=2.0, mel_fb=bank, out_kind=L.OUT_MEL)
    elif args.mfcc:
        plan = eng.plan(out_kind=L.OUT_MFCC, dct=T.dct_matrix(40, 80, "ortho"), **kw)
    else:
        plan = eng.plan(out_kind=L.OUT_LOGMEL, **kw)
    g = torch.Generator(device="cuda").manual_seed(4)
    wave = torch.clamp(0.05 * torch.randn(args.batch * n, generator=g, device="cuda"), -1, 1)
    batch = eng.batch(plan, np.arange(args.batch + 1, dtype=np.int64) * n)
    frames = batch.total_frames
    out = torch.empty((frames, plan.out_dim), dtype=torch.float32, device="cuda")

    def step():
        L.check(eng.lib.mafe_frontend_run(eng.ctx, plan.h, batch.h, C.c_void_p(wave.data_ptr()), L.WAVE_F32, 1.0,
                                          C.c_void_p(out.data_ptr()), L.DBGROUP_BATCH))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    eng.profile(True)
    eng.profile_reset()
    for _ in range(args.steps):
        step()
    k_ms, k_n = eng.profile_read(L.PROF_FBANK_MAIN)
    eng.profile(False)
    hours = args.batch * args.seconds / 3600.0
    bpf = (300 if args.fastspeech2 else 160) * 4 + plan.out_dim * 4
    what = ("melspectrogram n_fft=2048 win=1200 hop=300 128 mel @22.05 kHz (fastspeech2), [%d, %d] f32" % (args.batch, n)) if args.fastspeech2 \
        else "features.%s n_fft=400 hop=160 80 mel dB top_db=80 (batch floor), [%d, %d] f32" % ("mfcc(40)" if args.mfcc else "fbank", args.batch, n)
    print(json.dumps({"workload": what, "frames": frames, "fast_path": plan.is_fast,
                      "ms": ms, "main_kernel_ms": k_ms / max(k_n, 1), "audio_hours_per_s": hours / (ms / 1e3), "GBps_algorithmic": bpf * frames / ms / 1e6}))
    batch.close()


if __name__ == "__main__":
    main()
