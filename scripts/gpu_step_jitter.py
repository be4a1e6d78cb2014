"""Per-step time of the eager SimPO loop (one CUDA event pair per step) next to a CUDA-graph replay of the same step:
separates kernel time from host-side launch jitter."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead  # noqa: E402

dev = torch.device("cuda:0")
H = E = 4096
V, B, T, L = 16384, 64, 576, 1


class P:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(1)
head = FusedGenHead(P).to(dev).to(torch.bfloat16)
hidden = torch.randn(2 * B, L + T, H, device=dev).to(torch.bfloat16)
labels = torch.cat([torch.full((2 * B, L), -100, dtype=torch.long, device=dev), torch.randint(0, V, (2 * B, T), device=dev)], 1)
HP = dict(beta=10.0, gamma_beta_ratio=0.5, label_smoothing=0.0, sft_weight=0.0, loss_type="sigmoid")


def step():
    head.zero_grad(set_to_none=True)
    hh = hidden.detach().requires_grad_(True)
    out = head.simpo(hh, labels, image_span=(L - 1, L - 1 + T), **HP)
    out.loss.backward()
    return out.loss


for _ in range(3):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
ev[0].record()
for i in range(30):
    step()
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(30)]
print("EAGER per-step ms:", " ".join(f"{m:.1f}" for m in ms), flush=True)
print(f"EAGER mean {sum(ms) / len(ms):.3f} min {min(ms):.3f} max {max(ms):.3f}", flush=True)
try:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    g.replay()
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(30):
        g.replay()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(30)]
    print("GRAPH per-step ms:", " ".join(f"{m:.1f}" for m in ms), flush=True)
    print(f"GRAPH mean {sum(ms) / len(ms):.3f} min {min(ms):.3f} max {max(ms):.3f} loss {float(loss):.5f}", flush=True)
except Exception as ex:
    print("GRAPH failed:", repr(ex)[:300])
