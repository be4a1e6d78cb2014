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


def _world(group) -> int:
    if not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def _avg_op(group):
    """Sum in the collective, scale afterwards.  On NCCL this lets the tuner pick the in-switch NVLS algorithm for the
    268 MB dW2 message (NVSwitch multicast reduction, 24 channels); with ``ReduceOp.AVG`` it falls back to a 32-channel
    ring, whose CTAs take more from the GEMMs running beside it: 30.1 vs 30.6 ms per step at 8 GPUs.
    OSPO_HEAD_ALLREDUCE_OP=avg restores the averaging operator."""
    import os

    if dist.get_backend(group) == "nccl" and os.environ.get("OSPO_HEAD_ALLREDUCE_OP", "sum") == "avg":
        return dist.ReduceOp.AVG
    return dist.ReduceOp.SUM


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """in-place average of the flat gradient buffer over the data-parallel group (no-op for world size 1)"""
    world = _world(group)
    if world == 1:
        return flat
    op = _avg_op(group)
    dist.all_reduce(flat, op=op, group=group)
    if op == dist.ReduceOp.SUM:
        flat.mul_(1.0 / world)
    return flat


def staged_allreduce_mean_(flat: torch.Tensor, split: int, group, run_stage1, run_stage2, run_stage3=None):
    """Backward in stages with the exchange overlapped: ``run_stage1()`` fills ``flat[:split]`` (dW2, 80 %
    of the buffer) and its all-reduce starts asynchronously; ``run_stage2()`` fills the rest (db1, dW1; db2 is there
    already) and that all-reduce starts; ``run_stage3()`` (dX) runs beside it.  Returns the last stage's result.
    Same values as running the stages and then ``allreduce_mean_(flat)``."""
    world = _world(group)
    run_stage1()
    if world == 1:
        out = run_stage2()
        return run_stage3() if run_stage3 is not None else out
    op = _avg_op(group)
    head, tail = flat[:split], flat[split:]
    w1 = dist.all_reduce(head, op=op, group=group, async_op=True)
    out = run_stage2()
    w2 = dist.all_reduce(tail, op=op, group=group, async_op=True)
    if run_stage3 is not None:
        out = run_stage3()
    w1.wait()
    w2.wait()
    if op == dist.ReduceOp.SUM:
        flat.mul_(1.0 / world)
    return out
