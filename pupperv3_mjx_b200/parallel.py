"""Multi-GPU plumbing: envs shard independently across ranks (one process per GPU); the only collective of the
path is a sum all-reduce of the 24-float episode-metric accumulator over NCCL (gloo on CPU tests).
"""

from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.distributed as dist

from . import abi, prng


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block of envs owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_keys(seed: int, n_total: int, rank: int, world: int) -> np.ndarray:
    """``split(PRNGKey(seed), n_total)`` sliced for this rank, so results do not depend on the rank count."""
    lo, hi = shard_range(n_total, rank, world)
    return prng.split(prng.PRNGKey(seed), n_total)[lo:hi]


def allreduce_episode_totals(totals: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the episode accumulator over all ranks (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
    return totals


def episode_report(totals: torch.Tensor) -> Dict[str, float]:
    """Per-episode means from the (all-reduced) accumulator, named like Brax's ``episode_metrics``."""
    t = totals.detach().float().cpu().numpy()
    n = max(float(t[0]), 1.0)
    out = {"episodes": float(t[0]), "sum_reward": float(t[1]) / n, "length": float(t[2]) / n, "terminations": float(t[22])}
    for i, name in enumerate(abi.METRIC_NAMES):
        out[name] = float(t[3 + i]) / n
    return out
