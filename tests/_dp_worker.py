"""TEST INFRASTRUCTURE -- one rank of the multi-GPU gradient-exchange test (launched by torchrun from
tests/test_gpu_head.py::test_nccl_gradient_exchange_values, one process per GPU, NCCL).

Checks, with real NCCL on real GPUs, the path ``FusedGenHead.simpo(process_group=...)`` takes
(ospo/utils/train.py:26-28 semantics: per-rank ``losses.mean()``, gradients averaged over ranks):
  * the exchanged flat gradient is bit-identical on all ranks and equals the mean of the local ones,
  * dX (which stays local) is bit-identical to the run without a group,
  * the mean of the shard gradients equals the single-GPU full-batch gradient,
  * the same through the overlapped (staged) and the single-collective form.
Writes a JSON verdict to argv[1] (rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import head_oracle as O          # noqa: E402  (test infrastructure: input generators only)
from ospo_b200 import FusedGenHead           # noqa: E402
from ospo_b200 import dist as D              # noqa: E402


def main():
    out_path = sys.argv[1]
    H, E, V, T, L = (int(v) for v in sys.argv[2:7])
    pairs_per_rank = int(sys.argv[7])
    mode = sys.argv[8] if len(sys.argv) > 8 else "p2p"      # p2p: NVLink peer-memory exchange; nccl: NCCL all-reduce
    os.environ["OSPO_HEAD_DP"] = mode
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    Bg = pairs_per_rank * world
    head32 = O.make_head(H, E, V, seed=900, w2_gain=2.0)

    class P:
        n_embed, image_token_embed, image_token_size = H, E, V

    fh = FusedGenHead(P)
    fh.load_state_dict(head32.state_dict(), strict=True)
    fh = fh.to(dev).to(torch.bfloat16)
    hc, hr, lc, lr = O.synthetic_simpo_batch(Bg, T, L, H, V, seed=901, dtype=torch.bfloat16)
    hidden_g, labels_g = torch.cat([hc, hr]).to(dev), torch.cat([lc, lr]).to(dev)
    hidden, labels = D.shard_concatenated(hidden_g, labels_g, rank, world)
    hp = dict(beta=10.0, gamma_beta_ratio=0.5, loss_type="sigmoid")
    span = (L - 1, L - 1 + T)

    def run(h, lab, g):
        fh.zero_grad(set_to_none=True)
        x = h.detach().clone().requires_grad_(True)
        o = fh.simpo(x, lab, image_span=span, process_group=g, **hp)
        o.loss.backward()
        torch.cuda.synchronize()
        return o, x.grad.clone(), fh._flat.clone()

    res = {"mode": mode}
    _, dx_loc, flat_loc = run(hidden, labels, None)
    for ov in ("1", "0"):                         # NCCL path: staged (overlapped) and single all-reduce
        os.environ["OSPO_HEAD_OVERLAP"] = ov
        _, dx_red, flat_red = run(hidden, labels, group)
        chk = D.dp_check(flat_loc, flat_red, group)
        chk["dx_local_bit_identical"] = bool(torch.equal(dx_red, dx_loc))
        w2g = fh.vision_head.weight.grad.float()
        chk["param_grad_matches_flat"] = bool(torch.equal(w2g, flat_red[:V * E].view(V, E).to(torch.bfloat16).float()))
        # run to run: the exchange adds the contributions in a fixed order
        _, _, flat_again = run(hidden, labels, group)
        chk["bit_reproducible"] = bool(torch.equal(flat_again, flat_red))
        res[f"overlap_{ov}"] = chk
    ex = fh._peer_exchange(group)
    res["peer_exchange_active"] = ex is not None
    res["multicast"] = bool(ex is not None and ex.multicast)
    if mode == "p2p" and ex is None:
        res["overlap_1"]["status"] = "peer exchange expected but not active"
    # full batch on one GPU: DDP's average of per-rank means == the full-batch mean for equal shards
    _, _, flat_full = run(hidden_g, labels_g, None)
    rel = float((flat_red.double() - flat_full.double()).norm() / flat_full.double().norm())
    res["vs_full_batch_rel_fro"] = rel
    ok = all(v["status"] == "ok" and v["dx_local_bit_identical"] and v["param_grad_matches_flat"]
             and (v["bit_reproducible"] or mode == "nccl")
             for k, v in res.items() if k.startswith("overlap")) and rel < 2e-3
    res["ok"] = ok
    flags = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res["ok_all_ranks"] = bool(int(flags) == 1)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
