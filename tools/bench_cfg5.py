"""BASELINE.json configs[4]: 1 M synthetic utterances END TO END (pinned H2D + fbank + utterance CMVN + D2H), utterance-sharded
over the GPUs of one box, beside (a) the host <-> device copy ceiling of the same bytes measured in the same run and (b) the
reference's CPU pipeline on the host cores (`mp.Pool`, examples/conformer/dataset.py:449,479; bench.py --impl reference).

    python tools/bench_cfg5.py [--utts 1000000] [--chunk 8192]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_cfg5.py

Every rank owns utts / N utterances (lengths: the cfg3 distribution, 1-20 s uniform, seeded per rank) and streams them
through `mafe_frontend_run_host` -- the host-facing call of the path -- in calls of `--chunk` utterances from a RING of
pinned PCM16 buffers (the format `read()` decodes, mindaudio/data/io.py:741-745; the values are the int16-scaled samples
of `read() * (1 << 15)`); the features land in a ring of pinned float32 buffers.  1 M utterances are 336 GB of PCM16: the
ring holds `--pool` distinct seeded chunks that are cycled (their bytes are really copied every time; synthesising 336 GB
of noise on the host would measure numpy, not the path).

Ceiling: the same H2D and D2H byte counts per call moved with plain pinned `cudaMemcpyAsync` on two streams (full duplex),
all ranks at once, no kernels -- what the host's memory / PCIe path gives this process layout.  `frac_of_copy_ceiling`
says how much of it the pipeline keeps; the scaling of the ceiling itself over N says whether the host is the wall.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mindaudio_b200 as ma  # noqa: E402
from mindaudio_b200 import _lib as L  # noqa: E402

SR, N_MELS = 16000, 80


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1000000, help="utterances over ALL ranks")
    ap.add_argument("--chunk", type=int, default=8192, help="utterances per host call")
    ap.add_argument("--pool", type=int, default=2, help="distinct pinned chunks in the ring")
    ap.add_argument("--sub-chunk", type=int, default=128, help="utterances per pipelined device chunk inside a call")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        try:   # one contiguous slice of the allowed CPUs per rank (first-touch of the pinned rings on "its" memory)
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, cpus[local * per:(local + 1) * per] or cpus)
        except (AttributeError, OSError):
            pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ["MAFE_DEVICE"] = str(local)
    pipe = ma.FbankPipeline(cmvn="utt", mean_norm=True, std_norm=True)
    assert pipe.plan.is_fast

    n_rank = args.utts // world
    n_calls = max(1, (n_rank + args.chunk - 1) // args.chunk)
    # ---- the ring: `pool` seeded chunks of `chunk` utterances, PCM16 in pinned memory ----
    ring = []
    for k in range(args.pool):
        rng = np.random.default_rng(5 + 7919 * rank + 104729 * k)
        lens = rng.integers(16000, 320001, size=args.chunk).astype(np.int64)
        so = np.zeros(args.chunk + 1, dtype=np.int64)
        np.cumsum(lens, out=so[1:])
        total = int(so[-1])
        g = torch.Generator(device=dev)
        g.manual_seed(5 + 7919 * rank + 104729 * k)
        h = torch.empty(total, dtype=torch.int16, pin_memory=True)
        for s in range(0, total, 1 << 26):
            e = min(total, s + (1 << 26))
            h[s:e].copy_(torch.round(torch.clamp(0.05 * torch.randn(e - s, generator=g, device=dev), -1.0, 1.0) * 32768.0).to(torch.int16))
        frames = int(sum(pipe.plan.num_frames(int(n)) for n in lens))
        o = torch.empty((frames, N_MELS), dtype=torch.float32, pin_memory=True)
        ring.append((so, h, o, total, frames))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- end to end: n_calls host calls, wall clock around them (the calls are synchronous) ----
    so, h, o, _, _ = ring[0]
    pipe.run_host(h.data_ptr(), so, o.data_ptr(), L.WAVE_I16, 1.0, args.sub_chunk)    # warm-up (buffers, plans)
    barrier()
    t0 = time.perf_counter()
    audio_s, h2d, d2h, done = 0.0, 0, 0, 0
    for c in range(n_calls):
        so, h, o, total, frames = ring[c % args.pool]
        n_here = min(args.chunk, n_rank - done)
        if n_here < args.chunk:   # last, partial call: a prefix of the chunk
            so_c = so[: n_here + 1]
            total, frames = int(so_c[-1]), int(sum(pipe.plan.num_frames(int(n)) for n in np.diff(so_c)))
        else:
            so_c = so
        pipe.run_host(h.data_ptr(), so_c, o.data_ptr(), L.WAVE_I16, 1.0, args.sub_chunk)
        audio_s += total / SR
        h2d += total * 2
        d2h += frames * N_MELS * 4
        done += n_here
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    hours = sum_over_ranks(audio_s / 3600.0)
    h2d_all, d2h_all = sum_over_ranks(float(h2d)), sum_over_ranks(float(d2h))

    # ---- copy ceiling: the same bytes, plain pinned copies on two streams, all ranks at once ----
    so, h, o, total, frames = ring[0]
    d_w = torch.empty(total, dtype=torch.int16, device=dev)
    d_o = torch.empty((frames, N_MELS), dtype=torch.float32, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    n_copy = min(n_calls, 8)
    barrier()
    t0 = time.perf_counter()
    for c in range(n_copy):
        _, hh, oo, _, _ = ring[c % args.pool]
        with torch.cuda.stream(s_in):
            d_w[: hh.numel()].copy_(hh[: d_w.numel()], non_blocking=True)
        with torch.cuda.stream(s_out):
            oo[: d_o.shape[0]].copy_(d_o[: oo.shape[0]], non_blocking=True)
    torch.cuda.synchronize()
    copy_s = max_over_ranks(time.perf_counter() - t0)
    copy_bytes = sum_over_ranks(float(n_copy * (min(h.numel(), d_w.numel()) * 2 + min(o.shape[0], d_o.shape[0]) * N_MELS * 4)))
    # audio-hours the ceiling would carry: bytes per audio-hour of this workload
    bytes_per_hour = (h2d_all + d2h_all) / hours
    ceiling = copy_bytes / copy_s / bytes_per_hour

    if rank == 0:
        print(json.dumps({
            "workload": "cfg5: %d utterances (1-20 s, 16 kHz) end to end, %d rank(s) x %d host calls of %d utterances, PCM16 pinned "
                        "ring of %d chunk(s) -> mafe_frontend_run_host (sub-chunks of %d on 3 streams) -> pinned float32 features"
                        % (args.utts, world, n_calls, args.chunk, args.pool, args.sub_chunk),
            "n_gpus": world, "utterances": int(n_rank * world), "audio_hours": hours, "seconds": e2e_s,
            "value": hours / e2e_s, "unit": "audio-hours/s", "host_dtype": "int16 (PCM16) pinned",
            "h2d_bytes": h2d_all, "d2h_bytes": d2h_all, "host_device_gbs": (h2d_all + d2h_all) / e2e_s / 1e9,
            "copy_ceiling": {"value": ceiling, "unit": "audio-hours/s", "gbs": copy_bytes / copy_s / 1e9,
                             "what": "the same H2D + D2H bytes per call as plain pinned cudaMemcpyAsync on two streams, all ranks at once, no kernels"},
            "frac_of_copy_ceiling": (hours / e2e_s) / ceiling,
            "timing": "host wall clock around the synchronous host calls, max over ranks", "scaling": "weak per call, strong over the 1 M"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
