"""N > 1 host logic on CPU: gloo, world size 2.  Checks pair sharding and that the flat-gradient all-reduce
reproduces single-process full-batch gradients of the oracle (DDP averaging with per-rank losses.mean())."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ospo_b200 import ops
from ospo_b200.dist import allreduce_mean_, pair_shard, shard_concatenated, staged_allreduce_mean_


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import head_oracle as O

    H, E, V, B, T, L = 16, 24, 64, 4, 6, 2
    head = O.make_head(H, E, V, seed=1, w2_gain=3.0)
    hc, hr, lc, lr = O.synthetic_simpo_batch(B, T, L, H, V, seed=2)
    hidden, labels = torch.cat([hc, hr]), torch.cat([lc, lr])
    hp = dict(beta=10.0, gamma_beta_ratio=0.5, loss_type="sigmoid")
    h_loc, l_loc = shard_concatenated(hidden, labels, rank, world)
    b = h_loc.shape[0] // 2
    out = O.simpo_step(head, h_loc[:b], h_loc[b:], l_loc[:b], l_loc[b:], backward=True, **hp)
    # pack exactly like the library's flat buffer: dW2 | dW1 | db2 | db1
    flat = torch.cat([out["dW2"].reshape(-1), out["dW1"].reshape(-1), out["db2"], out["db1"]]).contiguous()
    assert flat.numel() == ops.flat_grad_numel(H, E, V)
    # the kernels store the weight gradients already times 1 / world (ospo_simpo_args.wgrad_scale): a plain sum then
    pre = flat / world
    allreduce_mean_(pre, dist.group.WORLD, prescaled=True)
    allreduce_mean_(flat, dist.group.WORLD)
    torch.testing.assert_close(pre, flat, rtol=1e-6, atol=1e-9)
    loss = out["loss"].detach().clone()
    dist.all_reduce(loss)
    if rank == 0:
        full = O.simpo_step(head, hc, hr, lc, lr, backward=True, **hp)
        dW2, dW1, db2, db1 = ops.split_flat_grads(flat, H, E, V)
        torch.testing.assert_close(dW2, full["dW2"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(dW1, full["dW1"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(db2, full["db2"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(db1, full["db1"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(loss / world, full["loss"].detach(), rtol=1e-6, atol=1e-7)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def _spawn_with_retry(worker, out_dir, attempts=2):
    """rendezvous on a fresh loopback port; one retry absorbs a port grabbed between probing and binding"""
    last = None
    for _ in range(attempts):
        try:
            mp.spawn(worker, args=(2, _free_port(), str(out_dir)), nprocs=2, join=True)
            return
        except Exception as ex:  # noqa: BLE001
            last = ex
    raise last


def test_pair_shard_layout():
    assert pair_shard(8, 1, 4) == slice(2, 4)
    with pytest.raises(ValueError):
        pair_shard(6, 0, 4)
    hidden = torch.arange(8).view(8, 1, 1).float()       # 4 chosen then 4 rejected
    labels = torch.arange(8).view(8, 1)
    h, l = shard_concatenated(hidden, labels, rank=1, world=2)
    assert h.flatten().tolist() == [2, 3, 6, 7] and l.flatten().tolist() == [2, 3, 6, 7]


def test_flat_gradient_allreduce_matches_full_batch_gloo_world2(tmp_path):
    _spawn_with_retry(_worker, tmp_path)
    assert (tmp_path / "ok").exists()


def _staged_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, split = 1000, 800
    g = torch.Generator().manual_seed(10 + rank)
    vals = torch.randn(n, generator=g)
    flat = torch.zeros(n)
    order = []

    def stage1():
        order.append(1)
        flat[:split] = vals[:split]

    def stage2():
        order.append(2)
        flat[split:] = vals[split:]

    def stage3():
        order.append(3)
        return "dx"

    out = staged_allreduce_mean_(flat, split, dist.group.WORLD, stage1, stage2, stage3)
    ref = vals.clone()
    allreduce_mean_(ref, dist.group.WORLD)
    assert out == "dx" and order == [1, 2, 3]
    torch.testing.assert_close(flat, ref, rtol=0, atol=0)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_staged_allreduce_equals_single_allreduce_gloo_world2(tmp_path):
    """the overlapped exchange (dW2 reduced while the second backward stage runs) gives the same flat buffer"""
    _spawn_with_retry(_staged_worker, tmp_path)
    assert (tmp_path / "ok").exists()
