"""BASELINE.json configs[3]: ECAPA-TDNN / fastspeech2 style front-end -- melspectrogram + mfcc(n_mfcc=40) on 3 s
utterances, global CMVN sufficient statistics all-reduced over the ranks (NCCL), then applied.  Device-resident,
one process per GPU:

    python tools/bench_cfg4.py [--utts 16384] [--steps 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_cfg4.py --utts 16384

One step per rank = for `--utts` utterances: melspectrogram (80 mel, n_fft 400, hop 160, power), mfcc (40 coefficients,
dB with the batch-wide top_db floor of features.py:263) -- by default from ONE transform (`mafe_frontend_run_aux`: the mel
energies are written beside the log-mel the DCT reads; `--two-transforms` = the two separate runs), Sigma x / Sigma x^2 / N of the MFCCs (float64), all-reduce of the
2*40+1 doubles, mean / istd (mindaudio/utils/load_files.py:19-28), (x - mean) * istd in place.  Weak scaling: every
rank owns its own utterances; the value is the audio of all ranks / the slowest rank's time.
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
import torch.distributed as dist  # noqa: E402

from mindaudio_b200 import _lib as L, _tables as T  # noqa: E402
from mindaudio_b200._engine import get_engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=16384, help="3 s utterances per rank per step")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--two-transforms", action="store_true",
                    help="melspectrogram and mfcc as two front-end runs (the reference's two calls; round-2 figure before "
                         "mafe_frontend_run_aux) instead of one run that writes the mel energies beside the MFCCs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine()
    eng.set_stream(torch.cuda.current_stream().cuda_stream or 1)

    n, D = 48000, 40
    bank = T.hz_triangle_bank(201, 80, 16000, 0.0, 8000.0)
    kw = dict(n_fft=400, hop=160, center=True, pad_mode="reflect", window=T.analysis_window("hann", 400, 400), power=2.0,
              mel_fb=bank)
    mel_plan = eng.plan(out_kind=L.OUT_MEL, **kw)
    mfcc_plan = eng.plan(out_kind=L.OUT_MFCC, dct=T.dct_matrix(D, 80, "ortho"), log_kind=L.LOG_DB, log_arg=1e-10,
                         log_mult=10.0, log_offset=0.0, top_db=80.0, **kw)
    g = torch.Generator(device="cuda").manual_seed(4 + rank)
    wave = torch.clamp(0.05 * torch.randn(args.utts * n, generator=g, device="cuda"), -1, 1)
    offs = np.arange(args.utts + 1, dtype=np.int64) * n
    b_mel, b_mfcc = eng.batch(mel_plan, offs), eng.batch(mfcc_plan, offs)
    frames = b_mel.total_frames
    mel = torch.empty((frames, 80), dtype=torch.float32, device="cuda")
    mfcc = torch.empty((frames, D), dtype=torch.float32, device="cuda")
    stats = torch.zeros(2 * D + 1, dtype=torch.float64, device="cuda")
    lib, ctx = eng.lib, eng.ctx
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def step():
        if args.two_transforms:
            L.check(lib.mafe_frontend_run(ctx, mel_plan.h, b_mel.h, vp(wave), L.WAVE_F32, 1.0, vp(mel), L.DBGROUP_NONE))
            L.check(lib.mafe_frontend_run(ctx, mfcc_plan.h, b_mfcc.h, vp(wave), L.WAVE_F32, 1.0, vp(mfcc), L.DBGROUP_BATCH))
        else:   # one transform: the n_fft 400 kernel writes the mel energies beside the dB log-mel the DCT reads
            L.check(lib.mafe_frontend_run_aux(ctx, mfcc_plan.h, b_mfcc.h, vp(wave), L.WAVE_F32, 1.0, vp(mfcc), L.DBGROUP_BATCH, vp(mel)))
        stats.zero_()
        L.check(lib.mafe_cmvn_stats_accumulate(ctx, vp(mfcc), frames, D, vp(stats)))
        if world > 1:
            dist.all_reduce(stats)                      # the one collective of the path: 81 doubles
        cnt = stats[2 * D]
        mean = stats[:D] / cnt
        var = torch.clamp(stats[D:2 * D] / cnt - mean * mean, min=1.0e-20)
        mean32, istd32 = mean.float().contiguous(), (1.0 / torch.sqrt(var)).float().contiguous()
        L.check(lib.mafe_cmvn_apply(ctx, vp(mfcc), frames, D, vp(mean32), vp(istd32)))
        return mean32

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        hours = world * args.utts * 3.0 / 3600.0
        print(json.dumps({"workload": "cfg4: melspectrogram(80) + mfcc(40, dB top_db=80 batch floor) + global CMVN "
                          "(stats all-reduced over %d rank(s), applied in place), %d x 3 s utterances per rank" % (world, args.utts),
                          "n_gpus": world, "frames_per_rank": frames, "ms_per_step": ms,
                          "transforms_per_step": 2 if args.two_transforms else 1,
                          "audio_hours_per_s": hours / (ms / 1e3), "fast_path": [mel_plan.is_fast, mfcc_plan.is_fast],
                          "cmvn_mean_head": [float(v) for v in m[:3].tolist()], "scaling": "weak",
                          "timing": "CUDA events, max over ranks"}))
    b_mel.close()
    b_mfcc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
