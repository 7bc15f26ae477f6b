"""Env sharding across GPUs: environments are independent, so each rank owns a contiguous block of
envs, weights are replicated and the rollout / reward / GAE stages need no data-path collective.  The only
cross-rank traffic of this path is bookkeeping (max-over-ranks timing, summed counters)."""

from __future__ import annotations


def env_shard(n_envs: int, rank: int, world: int) -> tuple[int, int]:
    """(start, count) of rank's contiguous env block; blocks differ by at most one env and cover [0, n_envs)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_envs, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def max_over_ranks(x: float, device=None) -> float:
    """MAX-reduce a host scalar over the default process group (gloo or nccl); identity when not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
