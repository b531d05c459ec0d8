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
