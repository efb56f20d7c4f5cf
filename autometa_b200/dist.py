"""Multi-GPU plumbing for the featurise path (one process per GPU, NCCL).

Contigs are sharded by total base pairs: counting, normalising
and projecting need no communication; the only exchange steps are tiny
all-reduces — the column-presence vector (am_clr's global column drop,
kmers.py:369), the column sums + row count (PCA mean) and the D x D Gram.
Everything here is device-agnostic torch.distributed code so that the N>1 host
logic is covered by world_size-2 ``gloo`` tests on CPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_ranges(lengths: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous contig ranges [lo, hi) per rank, balanced by total bp: rank r gets
    the contigs whose cumulative-bp midpoint falls in its 1/world_size slice."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    csum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        # first contig whose start is >= target (contigs are never split across GPUs)
        cuts.append(int(np.searchsorted(csum[:-1], target, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.asarray(cuts))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world_size)]


def is_distributed(group=None) -> bool:
    import torch.distributed as dist

    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allreduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of each tensor (Gram, column sums, row count)."""
    import torch.distributed as dist

    if not is_distributed(group):
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def allreduce_max_(t, group=None):
    """In-place MAX all-reduce (the 0/1 column-presence vector: logical OR)."""
    import torch.distributed as dist

    if is_distributed(group):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def allreduce_min_(t, group=None):
    import torch.distributed as dist

    if is_distributed(group):
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def allgather_rows(t, group=None):
    """Concatenate row shards of unequal length from all ranks (rank order)."""
    import torch
    import torch.distributed as dist

    if not is_distributed(group):
        return t
    world = dist.get_world_size(group)
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin this process (and the pages it first-touches from now on, e.g. pinned staging buffers) to the CPU
    cores NVML reports as local to the GPU.  With one process per GPU on a two-socket host this keeps every
    rank's host->device copies off the inter-socket link.  Returns False when NVML or the affinity call is
    unavailable (the run then continues unbound)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
