"""Soak test: many decode steps (eager on three streams round-robin + CUDA-graph replays) and SimPO steps back to back.
Checks determinism (same inputs -> same ids / loss every time), that the device flag words are re-armed after every
launch, that nothing trips the watchdog and that memory use is flat."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead, _abi  # noqa: E402

dev = torch.device("cuda:0")
H = E = 4096
V, P = 16384, 16


class Pm:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(0)
head = FusedGenHead(Pm).to(dev).to(torch.bfloat16)
for p in head.parameters():
    p.requires_grad_(True)
steps = 576
h = torch.randn(steps, 2 * P, H, device=dev).to(torch.bfloat16)
u = torch.rand(steps, P, device=dev)
ref = torch.empty(steps, P, dtype=torch.int64, device=dev)
with torch.no_grad():
    for i in range(steps):
        head.cfg_sample(h[i], 5.0, 1.0, uniforms=u[i], out=ref[i])
torch.cuda.synchronize()
t0 = time.time()
streams = [torch.cuda.Stream() for _ in range(3)]
outs = [torch.empty(steps, P, dtype=torch.int64, device=dev) for _ in range(3)]
for s in streams:
    s.wait_stream(torch.cuda.current_stream())
with torch.no_grad():
    for rep in range(2):
        for i in range(steps):
            for k, s in enumerate(streams):
                with torch.cuda.stream(s):
                    head.cfg_sample(h[i], 5.0, 1.0, uniforms=u[i], out=outs[k][i])
torch.cuda.synchronize()
for k in range(3):
    assert torch.equal(outs[k], ref), f"stream {k} diverged"
print(f"SOAK eager 3 streams x 2 x {steps} decode steps ok ({time.time() - t0:.1f} s)", flush=True)
g = torch.cuda.CUDAGraph()
gout = torch.empty(steps, P, dtype=torch.int64, device=dev)
with torch.no_grad():
    with torch.cuda.graph(g):
        for i in range(steps):
            head.cfg_sample(h[i], 5.0, 1.0, uniforms=u[i], out=gout[i])
for rep in range(20):
    gout.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(gout, ref), f"graph replay {rep} diverged"
print("SOAK 20 graph replays of 576 decode steps ok", flush=True)
# SimPO steps interleaved with decode steps
B, T, L = 16, 576, 1
hidden = torch.randn(2 * B, L + T, H, device=dev).to(torch.bfloat16)
labels = torch.cat([torch.full((2 * B, L), -100, dtype=torch.long, device=dev), torch.randint(0, V, (2 * B, T), device=dev)], 1)
mem0 = None
losses = []
for it in range(60):
    head.zero_grad(set_to_none=True)
    x = hidden.detach().requires_grad_(True)
    out = head.simpo(x, labels, image_span=(L - 1, L - 1 + T), beta=10.0, gamma_beta_ratio=0.5)
    out.loss.backward()
    with torch.no_grad():
        head.cfg_sample(h[it], 5.0, 1.0, uniforms=u[it], out=gout[it])
    losses.append(out.loss.detach())
    if it == 5:
        torch.cuda.synchronize()
        mem0 = torch.cuda.memory_allocated()
torch.cuda.synchronize()
ls = torch.stack(losses)
assert bool((ls == ls[0]).all()), "SimPO loss not reproducible"
assert torch.equal(gout[:60], ref[:60])
assert torch.isfinite(head.vision_head.weight.grad).all()
assert torch.cuda.memory_allocated() <= mem0 + (64 << 20), (torch.cuda.memory_allocated(), mem0)
wd = _abi.watchdog_record()
assert wd is None or wd[0] == 0, wd
print(f"SOAK 60 SimPO steps interleaved with decode ok: loss {float(ls[0]):.6f}, memory flat, watchdog clean", flush=True)
