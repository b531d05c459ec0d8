"""
Env sharding across ranks (one process per GPU).  The step path has no collective: envs are independent,
rank r owns the global envs [r*N, (r+1)*N), candle tables are replicated, and env i trades pair (i % num_pairs) where i
is the GLOBAL id -- so N per rank must be a multiple of num_pairs for shards to line up with a single-process run.
torch.distributed is used only for the barrier around the timed region and the max-over-ranks of its duration.
"""
from __future__ import annotations

import numpy as np

from .synth import start_offsets


def shard_range(envs_per_rank: int, rank: int, world: int) -> range:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return range(rank * envs_per_rank, (rank + 1) * envs_per_rank)


def shard_starts(envs_per_rank: int, rank: int, world: int, T: int, steps: int, S: int) -> np.ndarray:
    """Start bars of this rank's envs: the slice of the global deterministic spread."""
    r = shard_range(envs_per_rank, rank, world)
    return start_offsets(envs_per_rank * world, T, steps, S)[r.start:r.stop]


def check_pair_alignment(envs_per_rank: int, num_pairs: int) -> None:
    if envs_per_rank % num_pairs:
        raise ValueError("envs per rank must be a multiple of num_pairs so that (global id % num_pairs) == "
                         "(local id % num_pairs) on every rank")


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Whole-job duration = the slowest rank's (timing contract of bench.py)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- learner-side collectives: only a PPO learner built on top of the env needs these; the step path has none

def allreduce_mean_grads(params, dist=None, group=None) -> None:
    """Average the gradients of `params` across ranks with ONE all-reduce of a flat bucket (the policy is tiny --
    ~0.3 M parameters -- so the collective is latency-bound: one bucket, not one call per tensor)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    import torch

    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def global_mean_std(x, dist=None, group=None, eps: float = 1e-8):
    """Mean / std of `x` over ALL ranks' elements from one all-reduce of [sum, sum of squares, count] (advantage
    normalisation of a sharded rollout).  Returns (mean, std) as 0-d tensors on x's device."""
    import torch

    x64 = x.reshape(-1).to(torch.float64)
    s = torch.stack([x64.sum(), (x64 * x64).sum(), torch.tensor(float(x64.numel()), dtype=torch.float64, device=x.device)])
    if dist is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    mean = s[0] / s[2]
    var = torch.clamp(s[1] / s[2] - mean * mean, min=0.0)
    return mean.to(x.dtype), torch.sqrt(var).to(x.dtype) + eps
