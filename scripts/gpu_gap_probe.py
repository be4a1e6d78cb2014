"""Where does step time outside the library's kernel spans come from?  Runs the bench's SimPO step in phases
(kernel spans on/off, nvidia-smi sampler on/off, cold/warm) with per-step CUDA events and a forward/backward split."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import ClockSampler  # noqa: E402
from ospo_b200 import FusedGenHead, _abi  # noqa: E402

dev = torch.device("cuda:0")
H = E = 4096
V, B, T, L = 16384, 64, 576, 1
N = 30


class P:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(1)
head = FusedGenHead(P).to(dev).to(torch.bfloat16)
hidden = torch.randn(2 * B, L + T, H, device=dev).to(torch.bfloat16)
labels = torch.cat([torch.full((2 * B, L), -100, dtype=torch.long, device=dev), torch.randint(0, V, (2 * B, T), device=dev)], 1)
HP = dict(beta=10.0, gamma_beta_ratio=0.5, label_smoothing=0.0, sft_weight=0.0, loss_type="sigmoid")
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(N + 1)]


class NvmlSampler:
    """the same three readings through NVML in a Python thread (no nvidia-smi process)"""

    def __init__(self):
        import threading

        import pynvml
        self.n = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.rows, self.stop_flag = [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def start(self):
        self.t.start()

    def _run(self):
        import time
        while not self.stop_flag:
            self.rows.append((self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM),
                              self.n.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                              self.n.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            time.sleep(0.1)

    def stop(self):
        self.stop_flag = True
        self.t.join()
        sm = sorted(r[0] for r in self.rows)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "power_w_max": max((r[1] for r in self.rows), default=None),
                "samples": len(sm)}


def phase(name, spans, sampler):
    _abi.profile_enable(spans)
    if spans:
        _abi.profile_read()
    cs = (NvmlSampler() if sampler == 2 else ClockSampler(0)) if sampler else None
    if cs:
        cs.start()
    torch.cuda.synchronize()
    ev[0][0].record()
    for i in range(N):
        head.zero_grad(set_to_none=True)
        hh = hidden.detach().requires_grad_(True)
        out = head.simpo(hh, labels, image_span=(L - 1, L - 1 + T), **HP)
        ev[i][1].record()
        out.loss.backward()
        ev[i][2].record()
        ev[i + 1][0].record()
    torch.cuda.synchronize()
    clk = cs.stop() if cs else None
    st = [ev[i][0].elapsed_time(ev[i + 1][0]) for i in range(N)]
    fw = [ev[i][0].elapsed_time(ev[i][1]) for i in range(N)]
    bw = [ev[i][1].elapsed_time(ev[i][2]) for i in range(N)]
    msg = (f"{name:28s} step mean {sum(st) / N:6.2f} min {min(st):6.2f} max {max(st):6.2f} | fwd {sum(fw) / N:5.2f} "
           f"bwd {sum(bw) / N:5.2f}")
    if spans:
        prof = _abi.profile_read()
        fwd_k = ("gemm1_bias_gelu", "gemm2_logits_lse", "scalar_stage")
        sf = sum(v[0] for k, v in prof.items() if k in fwd_k) / N
        sb = sum(v[0] for k, v in prof.items() if k not in fwd_k) / N
        msg += f" | spans fwd {sf:5.2f} bwd {sb:5.2f} sum {sf + sb:6.2f}"
    if clk:
        msg += f" | sm {clk['sm_mhz']} W {clk['power_w_max']} n {clk['samples']}"
    print(msg, flush=True)
    print("   first 10 steps:", " ".join(f"{m:.1f}" for m in st[:10]), flush=True)
    _abi.profile_enable(False)


for _ in range(3):
    head.zero_grad(set_to_none=True)
    hh = hidden.detach().requires_grad_(True)
    head.simpo(hh, labels, image_span=(L - 1, L - 1 + T), **HP).loss.backward()
phase("cold  spans=0 smi=0", False, False)
phase("      spans=1 smi=0", True, False)
phase("      spans=1 smi=1", True, True)
phase("      spans=0 smi=1", False, True)
phase("warm  spans=0 smi=0", False, False)
phase("      spans=1 smi=1 again", True, True)
phase("      spans=1 nvml thread", True, 2)
phase("      spans=0 smi=0 last", False, False)
phase("      spans=1 smi=1 last", True, True)
