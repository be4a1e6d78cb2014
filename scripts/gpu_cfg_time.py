"""Per-step time of the configs[3] decode loop (576 steps in one CUDA graph) for the decode variants."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead, _abi, ops  # noqa: E402

import os

dev = torch.device("cuda:0")
H = E = int(os.environ.get("HE", "4096"))
ALT = int(os.environ.get("ALT", "1"))  # 1: two weight copies alternate (every step streams from HBM)
V, P, steps = 16384, 16, 576


class Pm:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(0)
head = FusedGenHead(Pm).to(dev).to(torch.bfloat16)
p = head._kernel_params()
alt = type(p)(p.w1.clone(), p.b1.clone(), p.w2.clone(), p.b2.clone())
PACK = int(os.environ.get("PACK", "1"))
pk = {id(p): (ops.pack_weight_impl(p.w1), ops.pack_weight_impl(p.w2)), id(alt): (ops.pack_weight_impl(alt.w1), ops.pack_weight_impl(alt.w2))}
h = torch.randn(steps, 2 * P, H, device=dev).to(torch.bfloat16)
u = torch.rand(steps, P, device=dev)
ids_out = torch.empty(steps, P, dtype=torch.int64, device=dev)
lib = _abi.load()
step_bytes = 2 * (H * E + E * V) + 4 * (E + V) + 2 * 2 * P * H + 4 * P + 8 * P


def run(direct):
    for i in range(steps):
        w = p if (i & 1) == 0 or not ALT else alt
        if direct:
            ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i], None,
                                pk[id(w)] if PACK else None)
        else:
            ids, _ = ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0)
            ids_out[i].copy_(ids)


ref = None
for merged, ahead, direct in ((1, 0, True), (1, 12, True), (1, 16, True), (1, 24, True), (1, 32, True), (1, 16, False),
                              (0, 0, True)):
    if True:
        lib.ospo_head_set_decode_merged(merged)
        lib.ospo_head_set_decode_l2_ahead(ahead)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run(direct)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run(direct)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        us = best * 1e3 / steps
        if ref is None:
            ref = ids_out.clone()
        same = bool(torch.equal(ref, ids_out))
        print(f"TIME pack={PACK} HE={H} alt={ALT} merged={merged} l2_ahead={ahead:2d} direct_out={int(direct)}: {us:6.2f} us/step  {step_bytes / us / 1e3:7.1f} GB/s  ids_same={same}")
lib.ospo_head_set_decode_merged(1)
lib.ospo_head_set_decode_l2_ahead(16)
