"""How fast does the swap-AB decode GEMM stream weights?  variant 101 (BN=32, K/K) on W2-like shapes, timed
inside CUDA graphs of 32 back-to-back launches (alternating two weight copies to defeat L2)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import _abi  # noqa: E402

lib = _abi.load()
dev = torch.device("cuda:0")
N = 32


def bench(M, K, tag, variant=101, copies=2):
    A = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(copies)]
    B = torch.randn(N * (M // 128 if variant == 305 else 1), K, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev)

    def run():
        for i in range(32):
            a = A[i % len(A)]
            lib.ospo_head_gemm_debug(variant, a.data_ptr(), a.stride(0), B.data_ptr(), B.stride(0), out.data_ptr(),
                                     out.stride(0), M, N, K, torch.cuda.current_stream().cuda_stream)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 5 / 32
    print(f"STREAM {tag} v{variant}: M={M} K={K} us={us:.2f} GB/s={M * K * 2 / us / 1e3:.0f}", flush=True)


for v in (300, 301, 302, 303, 304, 305, 101, 106):
    bench(16384, 4096, "W2", v)
for v in (301, 304, 305):
    bench(16384, 2048, "W2(1B)", v)
for v in (300, 301, 302, 303):
    bench(4096, 4096, "W1", v)
# the same matrix every launch: 33.5 MB stay in L2 -- what one SM can take in from L2 (301: 32 CTAs x 128 rows ... 302 / 303: all SMs)
for v in (301, 302, 303):
    bench(4096, 4096, "W1 L2-resident", v, copies=1)
for v in (301, 302, 303):
    bench(16384, 2048, "W2(1B) L2-resident", v, copies=1)
