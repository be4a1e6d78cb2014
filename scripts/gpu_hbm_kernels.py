"""One launch of every HBM-bound side kernel at the 7B sizes (for an ncu metrics pass): clip + AdamW, squared norm,
db2 / db1 column sums (through a SimPO step), weight packing."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead, ops  # noqa: E402

dev = torch.device("cuda:0")
H = E = 4096
V, B, T, L = 16384, 16, 576, 1


class P:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(1)
head = FusedGenHead(P).to(dev).to(torch.bfloat16)
hidden = torch.randn(2 * B, L + T, H, device=dev).to(torch.bfloat16).requires_grad_(True)
labels = torch.cat([torch.full((2 * B, L), -100, dtype=torch.long, device=dev), torch.randint(0, V, (2 * B, T), device=dev)], 1)
head.simpo(hidden, labels, image_span=(L - 1, L - 1 + T), beta=10.0, gamma_beta_ratio=0.5).loss.backward()
n = ops.flat_grad_numel(H, E, V)
g = torch.randn(n, device=dev) * 1e-3
p = torch.randn(n, device=dev) * 0.02
m = torch.zeros(n, device=dev)
v = torch.zeros(n, device=dev)
shadow = torch.empty(V * E + E * H, dtype=torch.bfloat16, device=dev)
sq = ops.grad_sqnorm_impl(g)
ops.adamw_step_impl(g, p, m, v, 1, 4e-5, 0.9, 0.95, 1e-8, 0.0, 1.0, sq, shadow)
ops.pack_weight_impl(head._kernel_params().w2)
torch.cuda.synchronize()
print("ok")
