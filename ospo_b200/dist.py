"""Data-parallel plumbing for the SimPO head.

Pairs are batch-sharded: rank r owns pairs [r*B/n, (r+1)*B/n) -- the chosen AND the rejected sequence of a
pair stay on the same rank, so the pair margin needs no communication.  The only exchange step is one
all-reduce (sum, then 1/world) of the contiguous fp32 gradient buffer dW2|dW1|db2|db1, which reproduces DDP's
gradient averaging with per-rank ``losses.mean()`` (ospo/utils/train.py:26-28, ospo/wrapper/train.py:419).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def pair_shard(num_pairs: int, rank: int, world: int) -> slice:
    if num_pairs % world:
        raise ValueError(f"{num_pairs} pairs do not shard evenly over {world} ranks")
    per = num_pairs // world
    return slice(rank * per, (rank + 1) * per)


def shard_concatenated(hidden: torch.Tensor, labels: torch.Tensor, rank: int, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """hidden/labels in the reference's concatenated layout [2B, ...] (chosen rows then rejected rows,
    train.py:364-365) -> this rank's [2B/world, ...] in the same layout"""
    B = hidden.shape[0] // 2
    s = pair_shard(B, rank, world)
    idx = torch.cat([torch.arange(s.start, s.stop), torch.arange(B + s.start, B + s.stop)]).to(hidden.device)
    return hidden.index_select(0, idx), labels.index_select(0, idx)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """in-place average of the flat gradient buffer over the data-parallel group (no-op for world size 1)"""
    if not dist.is_available() or not dist.is_initialized():
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / world)
    return flat
