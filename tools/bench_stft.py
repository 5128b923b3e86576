"""Secondary measurement: BASELINE.json configs[1] -- spectrum.stft + magphase, n_fft=512 hop=256 hann, a batch of
1024 synthetic 16 kHz 10 s utterances on one B200, device-resident (CUDA events).  HBM-bound path:
3080 B/frame for the STFT (256 samples in, 257 complex64 out), +3084 B/frame with magnitude and phase written.

    python tools/bench_stft.py [--batch 1024] [--seconds 10] [--steps 10]
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
    ap.add_argument("--n-fft", type=int, default=512)
    ap.add_argument("--hop", type=int, default=256)
    args = ap.parse_args()
    eng = get_engine()
    eng.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    n = int(args.seconds * 16000)
    plan = eng.plan(n_fft=args.n_fft, hop=args.hop, center=True, pad_mode="constant", out_kind=L.OUT_COMPLEX,
                    window=T.analysis_window("hann", args.n_fft, args.n_fft))
    g = torch.Generator(device="cuda").manual_seed(2)
    wave = torch.clamp(0.05 * torch.randn(args.batch * n, generator=g, device="cuda"), -1, 1)
    batch = eng.batch(plan, np.arange(args.batch + 1, dtype=np.int64) * n)
    frames, nb = batch.total_frames, args.n_fft // 2 + 1
    spec = torch.empty((frames, 2 * nb), dtype=torch.float32, device="cuda")
    mag = torch.empty((frames, nb), dtype=torch.float32, device="cuda")
    phase = torch.empty((frames, 2 * nb), dtype=torch.float32, device="cuda")

    def stft():
        L.check(eng.lib.mafe_frontend_run(eng.ctx, plan.h, batch.h, C.c_void_p(wave.data_ptr()), L.WAVE_F32, 1.0,
                                          C.c_void_p(spec.data_ptr()), L.DBGROUP_NONE))

    def magphase():
        L.check(eng.lib.mafe_magphase(eng.ctx, C.c_void_p(spec.data_ptr()), frames * nb, 1.0, C.c_void_p(mag.data_ptr()),
                                      C.c_void_p(phase.data_ptr())))

    res = {}
    for name, fn, bytes_per_frame in (("stft", stft, args.hop * 4 + nb * 8), ("magphase", magphase, nb * 8 + nb * 4 + nb * 8)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[name] = {"ms": ms, "GBps": bytes_per_frame * frames / ms / 1e6, "bytes_per_frame": bytes_per_frame}
    hours = args.batch * args.seconds / 3600.0
    total_ms = res["stft"]["ms"] + res["magphase"]["ms"]
    peak = 6554.9
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    print(json.dumps({"workload": "cfg2: stft n_fft=%d hop=%d hann + magphase, [%d, %d] f32" % (args.n_fft, args.hop, args.batch, n),
                      "frames": frames, "fast_path": plan.is_fast, "stft": res["stft"], "magphase": res["magphase"],
                      "audio_hours_per_s_stft": hours / (res["stft"]["ms"] / 1e3),
                      "audio_hours_per_s_stft_magphase": hours / (total_ms / 1e3),
                      "hbm_peak_GBps": peak, "stft_frac_of_hbm": res["stft"]["GBps"] / peak,
                      "magphase_frac_of_hbm": res["magphase"]["GBps"] / peak}))
    batch.close()


if __name__ == "__main__":
    main()
