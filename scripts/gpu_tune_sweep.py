"""Tuning aid: one SimPO step (configs[1] shape) per rasterisation / L2-eviction-hint setting of the six training
GEMMs, meant to be run under ``ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
-k regex:gemm_kernel`` so that the DRAM bytes of every setting can be read off the launch list.  Prints the setting
of each step in launch order (6 GEMM launches per step + the empty repair launch)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ospo_b200 import FusedGenHead, _abi  # noqa: E402

H = E = 4096
V, B, T, L = 16384, 64, 576, 1
OPTIONS = {   # kernel -> list of (group_m, a_evict, b_evict)
    0: [(16, 0, 0), (8, 0, 0), (32, 0, 0), (64, 0, 0), (16, 1, 2), (32, 1, 2), (64, 1, 2), (16, 2, 1), (32, 2, 1), (64, 2, 1)],
    1: [(16, 0, 0), (8, 0, 0), (32, 0, 0), (64, 0, 0), (16, 2, 1), (32, 2, 1), (64, 2, 1), (16, 1, 2), (8, 1, 2), (24, 2, 1)],
    2: [(16, 0, 0), (8, 0, 0), (9, 0, 0), (12, 0, 0), (4, 0, 0), (8, 1, 2), (9, 1, 0), (16, 1, 2), (12, 1, 2), (6, 1, 2)],
    3: [(16, 0, 0), (8, 0, 0), (9, 0, 0), (12, 0, 0), (4, 0, 0), (8, 1, 1), (9, 1, 1), (6, 0, 0), (10, 0, 0), (8, 2, 1)],
    4: [(16, 0, 0), (8, 0, 0), (4, 0, 0), (2, 0, 0), (8, 1, 1), (4, 1, 1), (6, 0, 0), (3, 0, 0), (5, 0, 0), (16, 1, 1)],
    5: [(16, 0, 0), (8, 0, 0), (32, 0, 0), (64, 0, 0), (16, 1, 2), (32, 1, 2), (64, 1, 2), (8, 1, 2), (16, 0, 2), (32, 0, 2)],
}


def main():
    dev = torch.device("cuda:0")
    lib = _abi.load()
    torch.manual_seed(1235)

    class P:
        n_embed, image_token_embed, image_token_size = H, E, V

    head = FusedGenHead(P).to(dev).to(torch.bfloat16)
    gen = torch.Generator(device=dev).manual_seed(1236)
    hidden = torch.randn(2 * B, L + T, H, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ids = torch.randint(0, V, (2 * B, T), generator=gen, device=dev)
    labels = torch.cat([torch.full((2 * B, L), -100, dtype=torch.long, device=dev), ids], 1)
    n = max(len(v) for v in OPTIONS.values())
    for i in range(n):
        cfg = {}
        for k, opts in OPTIONS.items():
            gm, ae, be = opts[i % len(opts)]
            lib.ospo_head_set_kernel_tune(k, gm, ae, be)
            cfg[k] = (gm, ae, be)
        head.zero_grad(set_to_none=True)
        x = hidden.detach().requires_grad_(True)
        out = head.simpo(x, labels, beta=10.0, gamma_beta_ratio=0.5, image_span=(L - 1, L - 1 + T))
        out.loss.backward()
        torch.cuda.synchronize()
        print(json.dumps({"step": i, "cfg": cfg, "loss": float(out.loss.detach())}), flush=True)


if __name__ == "__main__":
    main()
