#!/usr/bin/env python
"""bench.py -- fbank audio-hours/sec (16 kHz, 80-mel) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] -- the conformer/deepspeech2 front-end:
Kaldi-like 80-mel fbank (25 ms / 10 ms, 512-point FFT; examples/conformer/dataset.py:117-168) +
per-utterance CMVN on a RAGGED batch of synthetic 16 kHz utterances, 1-20 s uniform (seed 3),
offsets array, no padding.  One step = one pass of the hot path over one chunk of
``--utts`` utterances per GPU (8192: SURVEY.md section 8d processes the 100k utterances in chunks
of <= 8192).  Weak scaling: every rank owns its own chunk; no data-path collective.

value   = audio-hours processed by all ranks / device time (inputs resident in HBM), CUDA events,
          max over ranks.  Inputs (5.5 GB/chunk) are far larger than L2, so no flush is needed.
e2e     = the same through the host-facing call: pinned host waveform -> H2D -> kernels -> D2H of
          the features, copies inside the timed region.  Primary = PCM16 staging (what read() decodes,
          mindaudio/data/io.py:741-745: lossless for the reference's own input); f32 staging beside it.
roofline= dominant kernel (fbank512_v6_kernel) timed with CUDA events on its own stream inside this
          run, algorithmic bytes (960 B/frame) and flops (14 253/frame) from SURVEY.md section 8d.
          `bound` names the BINDING roofline (FP32 FMA peak measured in this run vs the measured HBM
          copy bandwidth), `frac` = that roofline's time / kernel time; `step` covers all kernels
          of the step with their DRAM traffic (tile records + the persistent kernel with the frame-mean
          sums and the utterance CMVN fused; the pre-pass / CMVN-apply slots are 0 unless the
          MAFE_NO_FUSED_* switches are set).
oracle_check = after the timed region a seeded sample of utterances of the TIMED output is compared
          with the float64 oracle (the checker, outside every timed region).
sustained = the same step repeated for >= --sustain-s seconds with the clocks sampled over that window.
cmvn_allreduce (N > 1) = one global-CMVN statistics pass (compute_cmvn_stats.py:104-128) whose NCCL
          all-reduce is checked against an all_gather + host sum; rank 0 writes the reference's JSON.
cpu_baseline / --impl reference = the reference's algorithm (oracle port, it is pure python and
          cannot travel to the GPU box) on this host's cores, bounded sample of the same workload;
          one BLAS thread per pool worker (stated in `sample`), one-process figure beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

SR = 16000
FRAME_LEN, HOP, N_MELS = 400, 160, 80
BYTES_PER_FRAME = 960.0      # 160 samples * 4 B in + 80 * 4 B out
FLOP_PER_FRAME = 14253.0     # SURVEY.md 8d (incl. utterance CMVN)
FP32_PEAK_NOMINAL_TFLOPS = 74.5   # 148 SM * 128 lanes * 2 * 1.965 GHz (not in MEASURED_PEAKS.json)
METRIC = "fbank audio-hours/sec (16kHz, 80-mel)"
UNIT = "audio-hours/s"


def chunk_lengths(seed, n_utts):
    """cfg3 length distribution: uniform 1-20 s."""
    return np.random.default_rng(seed).integers(16000, 320001, size=n_utts).astype(np.int64)


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as exc:  # NVML missing: report nothing rather than fail the bench
            self.reasons.add("nvml_unavailable:%s" % type(exc).__name__)

    def result(self):
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons)}


def cpu_baseline(n_utts, seed=3, single_utts=0):
    from oracle import cpu_baseline as cb         # the labelled CPU arm: the ONLY use of oracle/ in a timed leg
    lens = chunk_lengths(seed, n_utts)
    r = cb.run_pool(lens, seed=seed)
    hours = r["audio_s"] / 3600.0
    out = {"value": hours / r["wall_s"], "unit": UNIT, "cores": r["procs"], "kind": "port",
           "sample": "%d utterances (%.1f audio-min) of the cfg3 workload, fbank + utterance CMVN, reference-structured "
                     "numpy port (oracle/cpu_baseline.py), multiprocessing.Pool(%d), %d BLAS/OpenMP thread per worker; "
                     "%.1f s wall, %.1f s CPU" % (n_utts, r["audio_s"] / 60.0, r["procs"], r["blas_threads_per_proc"],
                                                   r["wall_s"], r["cpu_s"]),
           "wall_s": r["wall_s"], "audio_hours": hours, "blas_threads_per_worker": r["blas_threads_per_proc"]}
    if single_utts > 0:   # the one-process figure BASELINE.md section 3 asks for beside the pool figure
        r1 = cb.run_single(lens[:single_utts], seed=seed)
        out["single_process"] = {"value": r1["audio_s"] / 3600.0 / r1["wall_s"], "unit": UNIT, "utterances": int(single_utts),
                                 "blas_threads": r1["blas_threads"], "wall_s": r1["wall_s"]}
    return out


def oracle_check(wave_dev, out_dev, lens, frame_offsets, n_check, seed=11):
    """Sampled check of the TIMED output against the float64 oracle (oracle/restated.py), outside every timed
    region: conformer fbank (examples/conformer/dataset.py:117-168) + utterance mean / std normalisation
    (examples/ECAPA-TDNN/spec_augment.py:43-70).  The north-star bound is on the log-mel values, so the
    normalised output is de-normalised with the oracle's own statistics first; reported: the largest mixed error
    |d| / max(1, |ref|) and the fraction of elements outside the PLAIN 1e-4 bound (pass: <= 1e-4 of them)."""
    from oracle import restated as R               # the checker, never the thing measured
    so = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=so[1:])
    rng = np.random.default_rng(seed)
    picks = sorted(set(int(i) for i in rng.choice(len(lens), size=min(n_check, len(lens)), replace=False)))
    worst, bad, total, worst_norm = 0.0, 0, 0, 0.0
    for u in picks:
        w = wave_dev[int(so[u]):int(so[u + 1])].cpu().numpy().astype(np.float64)
        got = out_dev[int(frame_offsets[u]):int(frame_offsets[u + 1])].cpu().numpy().astype(np.float64)
        ref = R.conformer_fbank(w)
        mu, sd = ref.mean(axis=0), ref.std(axis=0)
        err = np.abs(got * sd + mu - ref) / np.maximum(1.0, np.abs(ref))
        worst = max(worst, float(err.max()))
        bad += int((err > 1e-4).sum())
        total += err.size
        worst_norm = max(worst_norm, float(np.max(np.abs(got - (ref - mu) / sd))))
    frac = bad / max(total, 1)
    return {"utterances": picks, "elements": total, "max_mixed_err_logmel": worst, "frac_outside_1e-4": frac,
            "max_abs_err_normalised": worst_norm, "ok": bool(frac <= 1e-4),
            "criterion": "|d| <= 1e-4 * max(1, |ref|) on the de-normalised log-mel; pass when <= 1e-4 of the elements exceed it"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this host's cores."""
    if rank != 0:
        return
    n = args.cpu_utts
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(max(n // 4, os.cpu_count() or 1))
    single = cpu_baseline(max(os.cpu_count() or 1, 8), single_utts=max(4, n // 32)).get("single_process")
    t_total = 0.0
    for _ in range(args.steps):
        r = cpu_baseline(n)
        vals.append(r)
        t_total += r["wall_s"]
    hours = sum(v["audio_hours"] for v in vals)
    value = hours / t_total
    last = vals[-1]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(n, 1, "host cores"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"],
                             "blas_threads_per_worker": last["blas_threads_per_worker"], "single_process": single},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_utts, world, where):
    return {"workload": "cfg3: conformer front-end = Kaldi-like 80-mel fbank (25ms/10ms, FFT 512) + utterance CMVN, "
                        "ragged batch of synthetic 16 kHz utterances 1-20 s (seed 3), offsets array",
            "utts_per_step_per_gpu": int(n_utts), "ranks": world, "where": where,
            "l2_policy": "inputs (~5.5 GB per step) exceed L2 (126 MB); no flush needed",
            "sharding": "by utterance, no data-path collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=8192, help="utterances per step per GPU")
    ap.add_argument("--cpu-utts", type=int, default=2048, help="utterances in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk-utts", type=int, default=128, help="utterances per chunk of the pipelined host path")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="seconds of the sustained-clock window (0: skip)")
    ap.add_argument("--oracle-utts", type=int, default=4, help="utterances of the timed output checked against the oracle (0: skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # CPU baseline first: it forks worker processes, which must happen before CUDA is initialised
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.cpu_utts, single_utts=max(4, args.cpu_utts // 32))

    if world > 1:
        # one contiguous slice of the allowed CPUs per rank: pinned buffers are then first-touched on the memory of
        # the socket that (on HGX boards) also hosts this rank's GPU, instead of all ranks sharing one NUMA node
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, cpus[local * per:(local + 1) * per] or cpus)
        except (AttributeError, OSError):
            pass

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ["MAFE_DEVICE"] = str(local)

    import mindaudio_b200 as ma
    from mindaudio_b200 import _lib as L
    from mindaudio_b200._engine import get_engine
    eng = get_engine()
    pipe = ma.FbankPipeline(cmvn="utt", mean_norm=True, std_norm=True)
    if not pipe.plan.is_fast:
        raise RuntimeError("the specialised sm_100a kernel is not serving the headline plan")
    pipe.use_torch_stream()

    # ---- synthetic chunk of this rank (seeded), resident in HBM ----
    lens = chunk_lengths(3 + 7919 * rank, args.utts)
    total = int(lens.sum())
    gen = torch.Generator(device=dev)
    gen.manual_seed(3 + 7919 * rank)
    wave = torch.empty(total, dtype=torch.float32, device=dev)
    step_elems = 1 << 26
    for s in range(0, total, step_elems):   # x = round(clip(0.05 N(0,1), -1, 1) * 32768): int16-scaled PCM (dataset.py:389-390)
        e = min(total, s + step_elems)
        wave[s:e] = torch.round(torch.clamp(0.05 * torch.randn(e - s, generator=gen, device=dev), -1.0, 1.0) * 32768.0)
    batch = pipe.layout(lens)
    n_frames = batch.total_frames
    out = torch.empty((n_frames, N_MELS), dtype=torch.float32, device=dev)
    audio_hours = total / SR / 3600.0

    def step():
        pipe.run(wave.data_ptr(), batch, out.data_ptr(), L.WAVE_F32, 1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        hrs = torch.tensor([audio_hours], dtype=torch.float64, device=dev)
        dist.all_reduce(hrs, op=dist.ReduceOp.SUM)
        total_hours = float(hrs.item())
    else:
        total_hours = audio_hours
    value = total_hours * args.steps / (ms / 1e3)

    # ---- sampled oracle check of the TIMED output (outside the timed region; rank 0) ----
    ocheck = None
    if rank == 0 and args.oracle_utts > 0:
        ocheck = oracle_check(wave, out, lens, batch.frame_offsets, args.oracle_utts)

    # ---- sustained clocks: the same step for >= --sustain-s seconds ----
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(np.ceil(args.sustain_s * 1e3 / max(ms / args.steps, 1e-3))))
        barrier()
        s_sampler = ClockSampler(local)
        s_sampler.start()
        sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sv0.record()
        for _ in range(n_sus):
            step()
        sv1.record()
        barrier()
        s_ms = sv0.elapsed_time(sv1)
        s_sampler.stop_flag = True
        s_sampler.join(timeout=2.0)
        if world > 1:
            t = torch.tensor([s_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t.item())
        sustained = {"seconds": s_ms / 1e3, "steps": n_sus, "ms_per_step": s_ms / n_sus,
                     "value": total_hours * n_sus / (s_ms / 1e3), "unit": UNIT, "clocks": s_sampler.result()}

    # ---- roofline of the dominant kernel: CUDA events around every launch, same steps ----
    eng.profile(True)
    eng.profile_reset()
    for _ in range(args.steps):
        step()
    k_ms, k_n = eng.profile_read(L.PROF_FBANK_MAIN)
    p_ms, p_n = eng.profile_read(L.PROF_FRAME_MEAN)
    c_ms, c_n = eng.profile_read(L.PROF_CMVN)
    o_ms, o_n = eng.profile_read(L.PROF_OTHER)
    eng.profile(False)
    # kernel time per step = sum over the step's launches (one today); the roofline's unit is the step's worth of frames
    k_avg_s = k_ms / args.steps / 1e3
    hbm_peak, peak_src = measured_peaks()
    alg_bytes = BYTES_PER_FRAME * n_frames
    alg_flops = FLOP_PER_FRAME * n_frames
    achieved_gbs = alg_bytes / k_avg_s / 1e9
    achieved_tflops = alg_flops / k_avg_s / 1e12
    # roofline = the slower of bytes at the measured HBM bandwidth and flops at the FP32 peak
    t_hbm = alg_bytes / (hbm_peak * 1e9)
    fp32_peak = eng.fp32_fma_peak()          # measured on this box, this run (FFMA microbenchmark in libmafe)
    t_fp32 = alg_flops / (fp32_peak * 1e12)
    # DRAM traffic per launch from the committed `ncu --set full` captures of this very workload (profiles/)
    traffic, traffic_src, ncu_pipes, step_traffic = None, None, None, None
    for name in ("r02_step_traffic.json", "r01_fbank512_traffic.json"):
        tp = os.path.join(REPO, "profiles", name)
        if not os.path.isfile(tp):
            continue
        with open(tp) as fh:
            td = json.load(fh)
        main_k = td.get("kernels", {}).get("main", td)
        if int(main_k.get("frames_per_launch", -1)) != int(n_frames):
            continue
        traffic, traffic_src = main_k["traffic_bytes_per_launch"], "profiles/%s (dram__bytes_read + dram__bytes_write)" % name
        ncu_pipes = {k: main_k[k] for k in ("kernel", "issue_active_pct", "pipe_fma_cycles_active_pct", "l1tex_data_pipe_pct",
                                            "pipe_lsu_pct", "tensor_pipe_pct", "registers_per_thread", "grid", "block",
                                            "warp_instructions", "shared_wavefronts") if k in main_k}
        if "kernels" in td:
            step_traffic = float(sum(k["traffic_bytes_per_launch"] for k in td["kernels"].values()))
        break
    step_s = ms / args.steps / 1e3
    binding_fp32 = t_fp32 >= t_hbm
    hbm_view = {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                "peak_source": peak_src + " (MEASURED_PEAKS.json)" if peak_src == "measured" else "fallback 6650 GB/s"}
    fp32_view = {"achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak,
                 "nominal_peak": FP32_PEAK_NOMINAL_TFLOPS, "frac_of_nominal": achieved_tflops / FP32_PEAK_NOMINAL_TFLOPS,
                 "peak_source": "FFMA microbenchmark of this run (mafe_fp32_fma_peak); MEASURED_PEAKS.json has no FP32 figure; "
                                "FFMA2 (packed) reaches the same lane rate at half the issue slots (profiles/r02_fp32x2_peak.txt)"}
    top = fp32_view if binding_fp32 else hbm_view
    roofline = {"bound": "fp32" if binding_fp32 else "hbm",
                "bound_note": "binding = the slower of algorithmic flops at the measured FP32 FMA peak and algorithmic bytes at "
                              "the measured HBM copy bandwidth (north star); tensor cores are not used on this path",
                "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes, "algorithmic_flops": alg_flops,
                "kernel": "fbank512_v6_kernel", "kernel_ms": k_avg_s * 1e3, "frames_per_launch": int(n_frames),
                "hbm": hbm_view, "fp32": fp32_view,
                "frac_of_min_roofline": max(t_hbm, t_fp32) / k_avg_s,
                "ncu_capture": ncu_pipes,
                "step": {"ms": step_s * 1e3, "frac_of_min_roofline": max(t_hbm, t_fp32) / step_s,
                         "hbm_frac_algorithmic": alg_bytes / step_s / 1e9 / hbm_peak,
                         "traffic": step_traffic, "traffic_over_algorithmic": (step_traffic / alg_bytes) if step_traffic else None,
                         "kernels_ms": {"frame_mean_prepass": p_ms / args.steps, "tile_records": o_ms / args.steps,
                                        "fbank512_v6_kernel": k_ms / args.steps, "cmvn_utt_apply": c_ms / args.steps},
                         "launches_per_step": [p_n // args.steps, o_n // args.steps, k_n // args.steps, c_n // args.steps],
                         "note": "per-kernel times are CUDA-event brackets around each kernel's launches, summed per step"}}

    # ---- the one collective of the path, under the driver's eyes (N > 1): global-CMVN statistics ----
    cmvn_check = None
    if world > 1:
        from mindaudio_b200.data.cmvn import CmvnStats, save_cmvn_json
        n_sub = min(256, len(lens))                    # a sub-batch of this rank's utterances is enough for the check
        sub = pipe.layout(lens[:n_sub])
        stats_dev = torch.zeros(2 * N_MELS + 1, dtype=torch.float64, device=dev)
        pipe.run_global_stats(wave.data_ptr(), sub, out.data_ptr(), stats_dev.data_ptr(), L.WAVE_F32, 1.0)
        torch.cuda.synchronize()
        mine = stats_dev.cpu().numpy()
        sub.close()
        st = CmvnStats(N_MELS)
        st.add_raw(mine)
        st.allreduce()                                   # NCCL all-reduce of 2 * D + 1 float64 (compute_cmvn_stats.py:104-112)
        gathered = [torch.zeros_like(stats_dev) for _ in range(world)]
        dist.all_gather(gathered, stats_dev)             # the other way: gather every rank's raw statistics, sum on the host
        ref = np.sum([g.cpu().numpy() for g in gathered], axis=0)
        got = np.concatenate([st.mean_stat, st.var_stat, [float(st.frame_num)]])
        rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
        ok = torch.tensor([1 if rel <= 1e-12 else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        cmvn_check = {"ok": bool(ok.item()), "max_rel_diff_vs_gather_sum": rel, "frame_num": int(st.frame_num),
                      "values": 2 * N_MELS + 1, "utterances_per_rank": int(n_sub), "backend": dist.get_backend()}
        if rank == 0:
            import tempfile
            path = os.path.join(tempfile.gettempdir(), "mafe_global_cmvn_n%d.json" % world)
            save_cmvn_json(st, path)                     # compute_cmvn_stats.py:121-128 format
            with open(path) as fh:
                cmvn_check["json_keys"] = sorted(json.load(fh).keys())

    # ---- end to end through the host-facing call: pinned host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        so = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=so[1:])
        h_out = torch.empty((n_frames, N_MELS), dtype=torch.float32, pin_memory=True)

        def run_e2e(h_wave, dtype_code, bytes_per_sample):
            """mafe_frontend_run_host: chunked H2D / kernels / D2H on three streams; synchronous host call, so the
            host wall clock around it is the honest end-to-end time (device events cannot see three streams)."""
            pipe.run_host(h_wave.data_ptr(), so, h_out.data_ptr(), dtype_code, 1.0, args.chunk_utts)   # warm-up
            barrier()
            n_e2e = max(2, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                pipe.run_host(h_wave.data_ptr(), so, h_out.data_ptr(), dtype_code, 1.0, args.chunk_utts)
            torch.cuda.synchronize()
            e_ms = (time.perf_counter() - t0) * 1e3
            if world > 1:
                t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e_ms = float(t.item())
            return {"value": total_hours * n_e2e / (e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": int(total * bytes_per_sample + 8 * len(so)),
                    "d2h_bytes_per_step": int(n_frames * N_MELS * 4), "steps": n_e2e, "ms_per_step": e_ms / n_e2e,
                    "chunk_utts": args.chunk_utts, "timing": "host wall clock around the synchronous host-to-host call"}

        # primary: the waveforms staged as PCM16 -- what read() decodes (io.py:741-745), i.e. the reference pipeline's own
        # input format (`read() * (1 << 15)` of a 16-bit WAV is integer valued): half the H2D bytes, bit-identical features
        h_i16 = torch.empty(total, dtype=torch.int16, pin_memory=True)
        h_i16.copy_(wave.to(torch.int16))
        torch.cuda.synchronize()
        e2e = run_e2e(h_i16, L.WAVE_I16, 2)
        e2e["host_dtype"] = "int16 (PCM16) pinned"
        del h_i16
        h_wave = torch.empty(total, dtype=torch.float32, pin_memory=True)
        h_wave.copy_(wave)
        torch.cuda.synchronize()
        e2e_f32 = run_e2e(h_wave, L.WAVE_F32, 4)
        e2e["f32_staging"] = {k: e2e_f32[k] for k in ("value", "h2d_bytes_per_step", "ms_per_step")}
        del h_wave

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args.utts, world, "B200"),
                "audio_hours_per_step": total_hours, "frames_per_step_rank0": int(n_frames),
                "clocks": sampler.result(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "oracle_check": ocheck, "sustained": sustained,
                "cmvn_allreduce": cmvn_check}
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
