"""Multi-GPU plumbing of the feature path: one process per GPU (``torchrun``), utterances sharded
across ranks, and ONE collective -- the all-reduce of the global-CMVN sufficient statistics
(2*D+1 float64 values; examples/conformer/compute_cmvn_stats.py accumulates them in a single
process, :104-112).  Everything else is independent per utterance, so there is no data-path
collective.

``shard_utterances`` mirrors the intent of the reference's samplers (``indices[rank::group_size]``,
mindaudio/utils/distributed.py:24-25; ``batch[rank::group_size]``, conformer/dataset.py:553) but
balances by TOTAL SAMPLES with contiguous ranges, which keeps every rank's slice of the flat
waveform array contiguous.
"""
from __future__ import annotations

import numpy as np


def shard_utterances(lengths, rank, world):
    """Contiguous utterance range [lo, hi) of ``rank``: the prefix sum of the lengths is cut at
    k/world of the total, so ranks get (nearly) equal numbers of samples."""
    lengths = np.asarray(lengths, dtype=np.int64)
    if world <= 1:
        return 0, len(lengths)
    csum = np.concatenate([[0], np.cumsum(lengths)])
    total = csum[-1]
    cuts = [int(np.searchsorted(csum, total * k / world, side="left")) for k in range(world + 1)]
    cuts[0], cuts[-1] = 0, len(lengths)
    for k in range(1, world + 1):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return cuts[rank], cuts[rank + 1]


def allreduce_cmvn_stats(stats, group=None):
    """Sum ``CmvnStats`` over the ranks (NCCL on GPUs, gloo in CPU tests); returns ``stats``."""
    return stats.allreduce(group)


def allreduce_max(value, group=None):
    """Batch-wide ``top_db`` floor when ONE reference call is sharded over ranks: the clamp of
    ``amplitude_to_dB`` on a 3-D batch uses the max of the whole call (spectrum.py:81-86)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    from ._engine import get_engine
    dev = torch.device("cuda", get_engine().device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
