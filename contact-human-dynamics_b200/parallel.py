"""Multi-GPU plumbing of the phys-optim path: sequences are independent NLPs (the reference runs them as separate
processes, scripts/run_phys_mocap.py:80), so they shard embarrassingly -- one process per GPU, no collective inside
the solver, one gather of the fixed-size sampled solutions at the end.

Everything here is backend agnostic (`nccl` on the GPU box, `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_by_work(work: Sequence[float], world: int) -> List[List[int]]:
    """Deals sequence indices to ranks: sorted by estimated work (number of variables ~ iterations x size),
    largest first, round robin with reversal every other pass (snake order) so that ragged batches balance.
    Deterministic; every rank computes the same assignment."""
    order = sorted(range(len(work)), key=lambda i: (-float(work[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for k, idx in enumerate(order):
        r = k % world
        if (k // world) % 2 == 1:
            r = world - 1 - r
        shards[r].append(idx)
    for s in shards:
        s.sort()
    return shards


def pad_to(shards: List[List[int]]) -> int:
    """Per-rank slot count so that the gather has equal receive counts."""
    return max((len(s) for s in shards), default=0)


def gather_samples(local, world: int, group=None):
    """all_gather of a (slots, frames, stride) tensor -> (world*slots, frames, stride).  `local` is a torch tensor
    on the device the process group's backend expects (cuda for nccl, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        return local
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def unshard(gathered: np.ndarray, shards: List[List[int]], slots: int) -> np.ndarray:
    """Reorders gathered rows (rank major, `slots` rows per rank) back to the original sequence order."""
    n = sum(len(s) for s in shards)
    out = np.zeros((n,) + gathered.shape[1:], dtype=gathered.dtype)
    for r, s in enumerate(shards):
        for k, idx in enumerate(s):
            out[idx] = gathered[r * slots + k]
    return out
