"""Three launches of the supplied-logits merge + sample kernel at the bench size (576 steps x 16 pairs), for ncu."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import cfg_merge_sample  # noqa: E402

dev = torch.device("cuda:0")
steps, P, V = 576, 16, 16384
g = torch.Generator(device=dev).manual_seed(5)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0  # 3: most tiles have codes under the 2^-120 cut-off; 1: none
lg = (torch.randn(steps, 2 * P, V, generator=g, device=dev) * scale).to(torch.bfloat16)
u = torch.rand(steps, P, generator=g, device=dev)
for _ in range(3):
    ids = cfg_merge_sample(lg, 5.0, 1.0, uniforms=u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    cfg_merge_sample(lg, 5.0, 1.0, uniforms=u)
e1.record()
torch.cuda.synchronize()
print("scale", scale, "us per launch", e0.elapsed_time(e1) * 100, "checksum", int(ids.sum()))
