"""A few CFG decode steps at the configs[3] shape (for ncu captures)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead, ops  # noqa: E402

dev = torch.device("cuda:0")
H = E = 4096
V, P = 16384, 16


class Pm:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(0)
head = FusedGenHead(Pm).to(dev).to(torch.bfloat16)
p = head._kernel_params()
h = torch.randn(4, 2 * P, H, device=dev).to(torch.bfloat16)
u = torch.rand(4, P, device=dev)
for i in range(4):
    ids, _ = ops.cfg_sample_impl(h[i], p.w1, p.b1, p.w2, p.b2, 5.0, 1.0, u[i], False, 0)
torch.cuda.synchronize()
print("ids", ids.tolist())
