"""Sharding of independent utterances / streams over ranks (one process per GPU, no data-path collective).

The reference has no multi-device inference (infer.py:35 picks one device; its only multi-GPU code is training
DDP, train.py:39-46,87-88).  Utterances and streams share nothing but read-only weights, so a job is split into
contiguous blocks, every rank runs its block through its own `GTCRNMicro`, and results are gathered on the host
of rank 0.  torch.distributed is used for rendezvous, the timing barrier / max-over-ranks and the host-side gather
only -- never on the hot path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of `n_items` for `rank`: ceil(n/world) items per rank, last ranks may be short/empty."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request n={n_items} rank={rank} world={world}")
    per = (n_items + world - 1) // world
    start = min(n_items, rank * per)
    return start, min(n_items, start + per)


def shard_plan(n_items: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n_items, r, world) for r in range(world)]


def gather_to_rank0(local: torch.Tensor, n_items: int, group=None) -> Optional[torch.Tensor]:
    """Host-side gather of per-rank result blocks (CPU tensors, dim 0 = items) into one tensor on rank 0, in item order."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    start, stop = shard_range(n_items, rank, world)
    if local.shape[0] != stop - start:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, its shard is [{start},{stop})")
    per = (n_items + world - 1) // world
    padded = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype)
    padded[:local.shape[0]] = local
    bufs = [torch.zeros_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([bufs[r][:b - a] for r, (a, b) in enumerate(shard_plan(n_items, world))], dim=0)


def max_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    """Slowest rank's time (the number a multi-GPU throughput is computed from)."""
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
