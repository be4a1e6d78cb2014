"""Per-kernel breakdown of the CFG decode step (eager, CUDA-event spans) + sampler-only timing."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead, _abi, cfg_merge_sample, ops  # noqa: E402

dev = torch.device("cuda:0")
H = E = 4096
V, P, steps = 16384, 16, 64


class Pm:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(0)
head = FusedGenHead(Pm).to(dev).to(torch.bfloat16)
p = head._kernel_params()
alt = type(p)(p.w1.clone(), p.b1.clone(), p.w2.clone(), p.b2.clone())
h = torch.randn(steps, 2 * P, H, device=dev).to(torch.bfloat16)
u = torch.rand(steps, P, device=dev)


def run():
    for i in range(steps):
        w = p if (i & 1) == 0 else alt
        ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0)


run()
torch.cuda.synchronize()
_abi.profile_enable(True)
_abi.profile_read()
run()
torch.cuda.synchronize()
prof = _abi.profile_read()
_abi.profile_enable(False)
print("CFGPROBE", json.dumps({k: round(1e3 * v[0] / v[1], 2) for k, v in prof.items()}), "us per launch")
for Pn in (1, 16, 9216):
    lg = (torch.randn(Pn * 2, V, device=dev) * 3).to(torch.bfloat16)
    uu = torch.rand(Pn, device=dev)
    for _ in range(3):
        cfg_merge_sample(lg, 5.0, 1.0, uniforms=uu)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        cfg_merge_sample(lg, 5.0, 1.0, uniforms=uu)
    e1.record()
    torch.cuda.synchronize()
    print(f"CFGPROBE sampler pairs={Pn} us={1e3 * e0.elapsed_time(e1) / 20:.2f}")

# whole-step timing under a CUDA graph (64 steps per graph)
ids_out = torch.empty(steps, P, dtype=torch.int64, device=dev)


def run_graphable():
    for i in range(steps):
        w = p if (i & 1) == 0 else alt
        ids, _ = ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0)
        ids_out[i].copy_(ids)


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    run_graphable()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run_graphable()
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"CFGPROBE graph us_per_step={1e3 * e0.elapsed_time(e1) / 10 / steps:.2f}")
