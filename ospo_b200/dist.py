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
    """Sum in the collective, scale afterwards (or before: see ``prescaled``).  On NCCL this lets the tuner pick the
    in-switch NVLS algorithm for the 268 MB dW2 message (NVSwitch multicast reduction, 24 channels); with
    ``ReduceOp.AVG`` it falls back to a 32-channel ring, whose CTAs take more from the GEMMs running beside it:
    30.1 vs 30.6 ms per step at 8 GPUs (round 1)."""
    return dist.ReduceOp.SUM


def allreduce_mean_(flat: torch.Tensor, group=None, prescaled: bool = False) -> torch.Tensor:
    """in-place average of the flat gradient buffer over the data-parallel group (no-op for world size 1).
    ``prescaled``: every rank's buffer already carries the 1 / world factor (the weight-gradient kernels apply it as
    they store, ``ospo_simpo_args.wgrad_scale``), so a plain sum is the average and no extra pass over the 336 MB
    buffer is needed."""
    world = _world(group)
    if world == 1:
        return flat
    op = dist.ReduceOp.SUM if prescaled else _avg_op(group)
    dist.all_reduce(flat, op=op, group=group)
    if op == dist.ReduceOp.SUM and not prescaled:
        flat.mul_(1.0 / world)
    return flat


def staged_allreduce_mean_(flat: torch.Tensor, split: int, group, run_stage1, run_stage2, run_stage3=None,
                           prescaled: bool = False):
    """Backward in stages with the exchange overlapped: ``run_stage1()`` fills ``flat[:split]`` (dW2, 80 %
    of the buffer) and its all-reduce starts asynchronously; ``run_stage2()`` fills the rest (db1, dW1; db2 is there
    already) and that all-reduce starts; ``run_stage3()`` (dX) runs beside it.  Returns the last stage's result.
    Same values as running the stages and then ``allreduce_mean_(flat)``."""
    world = _world(group)
    run_stage1()
    if world == 1:
        out = run_stage2()
        return run_stage3() if run_stage3 is not None else out
    op = dist.ReduceOp.SUM if prescaled else _avg_op(group)
    head, tail = flat[:split], flat[split:]
    w1 = dist.all_reduce(head, op=op, group=group, async_op=True)
    out = run_stage2()
    w2 = dist.all_reduce(tail, op=op, group=group, async_op=True)
    if run_stage3 is not None:
        out = run_stage3()
    w1.wait()
    w2.wait()
    if op == dist.ReduceOp.SUM and not prescaled:
        flat.mul_(1.0 / world)
    return out


class PeerGradExchange:
    """Gradient exchange of the head's flat gradient over NVLink peer memory (``ospo_dp_exchange`` in
    include/ospo_head.h) -- DDP's averaging (ospo/utils/train.py:26-28) without a collective library on the data path.

    Two symmetric-memory buffers per rank (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every
    rank of the box): the flat gradient itself and an inbox of the same size.  Step 1 runs INSIDE the backward: the
    weight-gradient GEMM epilogues store every tile, times 1 / world, into the inbox of the rank that owns those rows
    (plain posted peer stores, no SM set aside, no extra kernel).  Step 2, after a barrier: each owner adds the `world`
    slots of its inbox in rank order and multicasts the sums into every rank's flat buffer (one ``multimem.st`` per 16
    bytes through the NVSwitch when a multicast mapping exists).  A second barrier and all ranks hold the same bits.
    torch.distributed is used for the rendezvous and the barriers only."""

    def __init__(self, group, H: int, E: int, V: int, device):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _abi, ops

        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not (2 <= self.world <= 8) or V % (self.world * 256) or E % (self.world * 256):
            raise ValueError("peer exchange needs 2..8 ranks and V, E divisible by 256 * world")
        n = ops.flat_grad_numel(H, E, V)
        self.dims = (H, E, V)
        name = group.group_name if hasattr(group, "group_name") else dist.group.WORLD.group_name
        self.flat = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.inbox = symm_mem.empty(n, dtype=torch.float32, device=device)      # world slots x (n / world) elements
        self.h_flat = symm_mem.rendezvous(self.flat, name)
        self.h_inbox = symm_mem.rendezvous(self.inbox, name)
        a = _abi.DpExchange()
        a.world, a.rank = self.world, self.rank
        for i in range(self.world):
            a.inbox[i] = int(self.h_inbox.buffer_ptrs[i])
            a.flat[i] = int(self.h_flat.buffer_ptrs[i])
        mc = int(getattr(self.h_flat, "multicast_ptr", 0) or 0)
        a.flat_multicast = mc if mc else None
        self.multicast = bool(mc)
        self.args = a
        self._shape = _abi.Shape(1, H, E, V, 1)
        self._C = C

    def barrier(self, channel: int = 0) -> None:
        self.h_inbox.barrier(channel=channel)

    def reduce_broadcast(self, regions: int = 3, max_blocks: int = 0) -> None:
        from . import _abi

        _abi.check(_abi.load().ospo_head_dp_reduce_broadcast(self._C.byref(self._shape), self._C.byref(self.args),
                                                            int(regions), int(max_blocks),
                                                            torch.cuda.current_stream().cuda_stream),
                   "ospo_head_dp_reduce_broadcast")

    def finish(self) -> None:
        """after the backward whose stores went to the inboxes: barrier, inbox -> everyone's flat, barrier"""
        self.barrier()
        self.reduce_broadcast()
        self.barrier()

    def run_staged(self, run_stage1, run_stage2, run_stage3):
        """Backward in three parts with both halves of the exchange overlapped.  ``run_stage1()`` ends with the dW2 GEMM,
        whose epilogue has scattered dW2 (80 % of the bytes) into the owners' inboxes: a side stream waits for every
        rank to get there, sums its dW2 shard and multicasts it while ``run_stage2()`` (db1, dW1) runs; the same for the
        remainder (dW1 rows, biases) while ``run_stage3()`` (dX) runs.  One barrier at the end: every rank's shards have
        landed in every flat buffer.  Returns run_stage3()'s result."""
        main = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream()
            self._side_blocks = 64   # a small grid shares the SMs with the GEMMs it runs beside (16 blocks: slower)
        run_stage1()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self.barrier(channel=1)
            self.reduce_broadcast(regions=1, max_blocks=self._side_blocks)
        run_stage2()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self.barrier(channel=1)
            self.reduce_broadcast(regions=2, max_blocks=self._side_blocks)
        out = run_stage3()
        main.wait_stream(self._side)
        self.barrier()
        return out


_peer_exchanges: dict = {}


def peer_exchange_for(group, H: int, E: int, V: int, device):
    """the PeerGradExchange of (group, head shape) when OSPO_HEAD_DP=p2p selects it, created on first use --
    collectively: every rank of the group must call this at the same point.  Returns None (and the caller uses the
    NCCL all-reduce, the default: on 8 B200s the two are within noise of each other, 30.8 vs 30.1 ms per step, see
    DESIGN.md §6) when the shape does not partition or symmetric memory cannot be set up on ANY rank."""
    import os

    key = (id(group), H, E, V, str(device))
    if key in _peer_exchanges:
        return _peer_exchanges[key]
    ex = None
    if os.environ.get("OSPO_HEAD_DP", "nccl") == "p2p" and dist.get_backend(group) == "nccl":
        ok = 1
        try:
            ex = PeerGradExchange(group, H, E, V, device)
        except Exception as e:  # noqa: BLE001 -- any failure means "not available here"; decided collectively below
            ok, ex = 0, None
            if os.environ.get("OSPO_HEAD_DP") == "p2p":
                print(f"[ospo_b200] peer-memory gradient exchange unavailable on rank {dist.get_rank(group)}: {e!r}",
                      flush=True)
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0:
            ex = None
    _peer_exchanges[key] = ex
    return ex


def dp_check(flat_local: torch.Tensor, flat_reduced: torch.Tensor, group=None) -> dict:
    """Value check of the gradient exchange on the hardware it runs on (DDP semantics, ospo/utils/train.py:26-28):
    ``flat_reduced`` (this rank's buffer after the exchange) must be bit-identical on every rank and equal the mean
    over ranks of the pre-exchange local buffers ``flat_local``.  Compares float64 checksums (sum and sum of squares)
    and a strided sample of elements gathered from all ranks.  Collective: every rank must call it."""
    world = _world(group)
    dev = flat_reduced.device
    idx = torch.arange(0, flat_reduced.numel(), max(1, flat_reduced.numel() // 65536), device=dev)
    loc = torch.cat([flat_local.double().sum().reshape(1), flat_local.double().pow(2).sum().reshape(1),
                     flat_local[idx].double()])
    red = torch.cat([flat_reduced.double().sum().reshape(1), flat_reduced.double().pow(2).sum().reshape(1),
                     flat_reduced[idx].double()])
    bits = flat_reduced.view(torch.int32).to(torch.int64)
    sig = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
    if world > 1:
        locs = [torch.empty_like(loc) for _ in range(world)]
        reds = [torch.empty_like(red) for _ in range(world)]
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(locs, loc, group=group)
        dist.all_gather(reds, red, group=group)
        dist.all_gather(sigs, sig, group=group)
    else:
        locs, reds, sigs = [loc], [red], [sig]
    mean_loc = torch.stack(locs).mean(0)
    ranks_equal = all(bool(torch.equal(s, sigs[0])) for s in sigs) and all(bool(torch.equal(r, reds[0])) for r in reds)
    scale = float(torch.stack(locs)[:, 2:].abs().max().clamp(min=1e-30))
    sample_err = float((reds[0][2:] - mean_loc[2:]).abs().max()) / scale
    sum_err = abs(float(reds[0][0] - mean_loc[0])) / max(float(torch.stack(locs)[:, 2:].abs().sum()), 1e-30)
    ok = ranks_equal and sample_err < 1e-5 and sum_err < 1e-5
    return {"status": "ok" if ok else "MISMATCH", "world": world, "ranks_bit_identical": ranks_equal,
            "reduced_checksum": float(reds[0][0]), "mean_of_local_checksums": float(mean_loc[0]),
            "sample_max_abs_err_rel": sample_err, "checksum_err_rel": sum_err, "sampled_elements": int(idx.numel())}
