"""N > 1 host logic on CPU: world_size-2 gloo.  Utterance sharding and the one collective of the
path (global-CMVN sufficient statistics, 2*D+1 float64) -- no GPU involved: the per-rank statistics
come from the oracle here, the all-reduce / finalisation code is the product's."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _feats(seed, t, d=8):
    return np.random.default_rng(seed).normal(10.0, 3.0, size=(t, d))


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mindaudio_b200.data.cmvn import CmvnStats
        from mindaudio_b200.dist import allreduce_cmvn_stats, allreduce_max, shard_utterances
        lo, hi = shard_utterances(lens, rank, world)
        st = CmvnStats(8)
        for u in range(lo, hi):                       # per-rank accumulation (device kernel on a GPU box)
            f = _feats(u, lens[u] // 100)
            st.add_raw(np.concatenate([f.sum(0), (f * f).sum(0), [f.shape[0]]]))
        allreduce_cmvn_stats(st)
        mx = allreduce_max(float(rank) - 3.5)
        q.put((rank, lo, hi, st.to_dict(), mx))
    finally:
        dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    lens = [int(v) for v in np.random.default_rng(3).integers(1600, 32000, size=37)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, d0, m0), (r1, lo1, hi1, d1, m1) = res
    assert (lo0, hi1) == (0, len(lens)) and hi0 == lo1                       # contiguous cover, no overlap
    s0, s1 = sum(lens[lo0:hi0]), sum(lens[lo1:hi1])
    assert abs(s0 - s1) <= max(lens)                                          # balanced by samples
    assert d0 == d1 and m0 == m1 == -2.5                                      # every rank holds the reduced values
    # equals the single-process accumulation of compute_cmvn_stats.py:104-112
    from oracle import restated as R
    n, a, b = R.cmvn_stats([_feats(u, lens[u] // 100) for u in range(len(lens))])
    assert d0["frame_num"] == n
    assert np.allclose(d0["mean_stat"], a, rtol=1e-12) and np.allclose(d0["var_stat"], b, rtol=1e-12)
    from mindaudio_b200.data.cmvn import cmvn_from_stats
    mean, istd = cmvn_from_stats(d0)
    rm, ri = R.cmvn_from_stats(n, a, b)
    assert np.allclose(mean, rm) and np.allclose(istd, ri)


def test_shard_edge_cases():
    from mindaudio_b200.dist import shard_utterances
    assert shard_utterances([5, 5, 5], 0, 1) == (0, 3)
    cuts = [shard_utterances([100] * 10, r, 4) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 10 and all(cuts[i][1] == cuts[i + 1][0] for i in range(3))
    cuts = [shard_utterances([7], r, 8) for r in range(8)]                    # fewer utterances than ranks
    assert sum(hi - lo for lo, hi in cuts) == 1
    assert [shard_utterances([], r, 2) for r in range(2)] == [(0, 0), (0, 0)]
