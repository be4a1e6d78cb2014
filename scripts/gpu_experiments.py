"""Ad-hoc GPU experiments for kernel tuning: per-kernel CUDA-event times of the SimPO head under different
settings (logits spill on/off, cta_group, group_m).  Output: gpurun_out/experiments.jsonl"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead, _abi, ops  # noqa: E402

dev = torch.device("cuda:0")
H = E = 4096
V, B, T = 16384, 64, 576
rows = 2 * B * T


class P:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(0)
head = FusedGenHead(P).to(dev).to(torch.bfloat16)
p = head._kernel_params()
x = torch.randn(rows, H, device=dev).to(torch.bfloat16)
labels = torch.randint(0, V, (rows,), device=dev)
seq_off = torch.arange(0, rows + 1, T, device=dev, dtype=torch.int64)
out = open(ROOT / "gpurun_out" / "experiments.jsonl", "a")
lib = _abi.load()


def emit(tag, prof):
    rec = {"tag": tag, **{k: round(v[0] / v[1], 4) for k, v in prof.items()}}
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    out.flush()


def fwd_bwd(save=True, need_dx=True, need_dw=True, iters=3):
    flat = torch.empty(ops.flat_grad_numel(H, E, V), device=dev) if need_dw else torch.empty(0, device=dev)
    one = torch.ones(1, device=dev)
    for it in range(iters + 1):
        if it == 1:
            torch.cuda.synchronize()
            _abi.profile_enable(True)
            _abi.profile_read()
        r = ops.simpo_fwd_impl(x, p.w1, p.b1, p.w2, p.b2, labels, seq_off, 10.0, 0.5, 0.0, 0.0, 0, save)
        if save:
            scalars, seq_logps, losses, crew, rrew, row_logps, row_lse, row_ref, grad_seq, pre, act, logits = r
            ops.head_bwd_impl(x, p.w1, p.b1, p.w2, p.b2, labels, seq_off, True, 0.0, scalars, pre, act, logits, row_lse,
                              row_ref, grad_seq, one, need_dx, flat, True)
    torch.cuda.synchronize()
    prof = _abi.profile_read()
    _abi.profile_enable(False)
    return prof


for cg in (2, 1):
    lib.ospo_head_set_cta_group(cg)
    for gm in (8, 16, 32):
        lib.ospo_head_set_group_m(gm)
        emit(f"cg{cg}_gm{gm}_full", fwd_bwd())
    lib.ospo_head_set_group_m(16)
    emit(f"cg{cg}_fwd_only_no_spill", fwd_bwd(save=False))
lib.ospo_head_set_cta_group(2)
emit("cg2_frozen_head_dx_only", fwd_bwd(need_dw=False))
