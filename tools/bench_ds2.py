"""Secondary measurement: the deepspeech2 front-end (examples/deepspeech2/dataset.py:39-47) --
stft(n_fft=320, hop=160, win=320, hann, centre/constant) -> magphase(power=1) -> log1p + scalar normalisation,
device-resident on `[batch, seconds * 16000]` float32.

    python tools/bench_ds2.py [--batch 1024] [--seconds 10] [--steps 10]
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
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--generic", action="store_true", help="force the generic kernel (A/B)")
    args = ap.parse_args()
    eng = get_engine()
    eng.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    n = int(args.seconds * 16000)
    plan = eng.plan(n_fft=320, hop=160, center=True, pad_mode="constant", out_kind=L.OUT_COMPLEX,
                    window=T.analysis_window("hann", 320, 320), allow_fast_path=not args.generic)
    g = torch.Generator(device="cuda").manual_seed(2)
    wave = torch.clamp(0.05 * torch.randn(args.batch * n, generator=g, device="cuda"), -1, 1)
    batch = eng.batch(plan, np.arange(args.batch + 1, dtype=np.int64) * n)
    frames = batch.total_frames
    spec = torch.empty((frames, 2 * 161), dtype=torch.float32, device="cuda")
    mag = torch.empty((frames, 161), dtype=torch.float32, device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def stft():
        L.check(eng.lib.mafe_frontend_run(eng.ctx, plan.h, batch.h, vp(wave), L.WAVE_F32, 1.0, vp(spec), L.DBGROUP_NONE))

    def rest():
        L.check(eng.lib.mafe_magphase(eng.ctx, vp(spec), frames * 161, 1.0, vp(mag), None))
        L.check(eng.lib.mafe_cmvn_scalar(eng.ctx, vp(mag), C.c_void_p(batch.frame_offsets_dev), args.batch, 161, 1))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    # round 2: the transform writes log1p(|X|) itself (MAFE_OUT_POWER, power 1, MAFE_LOG_LN_PLUS 1), then the scalar norm
    plan_f = eng.plan(n_fft=320, hop=160, center=True, pad_mode="constant", out_kind=L.OUT_POWER, power=1.0,
                      log_kind=L.LOG_LN_PLUS, log_arg=1.0, window=T.analysis_window("hann", 320, 320), allow_fast_path=not args.generic,
                      utt_scalar_norm=True)
    plan_t = eng.plan(n_fft=320, hop=160, center=True, pad_mode="constant", out_kind=L.OUT_POWER, power=1.0,
                      log_kind=L.LOG_LN_PLUS, log_arg=1.0, window=T.analysis_window("hann", 320, 320), allow_fast_path=not args.generic)
    batch_t = eng.batch(plan_t, np.arange(args.batch + 1, dtype=np.int64) * n)
    batch_f = eng.batch(plan_f, np.arange(args.batch + 1, dtype=np.int64) * n)

    def fused():
        L.check(eng.lib.mafe_frontend_run(eng.ctx, plan_f.h, batch_f.h, vp(wave), L.WAVE_F32, 1.0, vp(mag), L.DBGROUP_NONE))

    def fused_transform_only():
        L.check(eng.lib.mafe_frontend_run(eng.ctx, plan_t.h, batch_t.h, vp(wave), L.WAVE_F32, 1.0, vp(mag), L.DBGROUP_NONE))

    t_stft = timed(stft)
    t_all = timed(lambda: (stft(), rest()))
    t_fused = timed(fused)
    t_fused_tr = timed(fused_transform_only)
    hours = args.batch * args.seconds / 3600.0
    print(json.dumps({"workload": "deepspeech2 front-end: stft 320/160 hann + magphase + log1p + scalar norm, [%d, %d] f32" % (args.batch, n),
                      "frames": frames, "fast_path": plan.is_fast, "stft_ms": t_stft, "stft_GBps": frames * (160 * 4 + 161 * 8) / t_stft / 1e6,
                      "front_end_ms": t_all, "fused_front_end_ms": t_fused, "fused_transform_ms": t_fused_tr,
                      "audio_hours_per_s_fused_front_end": hours / (t_fused / 1e3), "audio_hours_per_s_stft": hours / (t_stft / 1e3),
                      "audio_hours_per_s_front_end": hours / (t_all / 1e3)}))
    batch.close()


if __name__ == "__main__":
    main()
