"""Single-wave DRAM traffic probe: 64 tiles (16 M-blocks x 4 N-blocks) on 74 CTA pairs -- does operand sharing
through L2 work when all clusters start together?  Run under ncu --metrics dram__bytes_read.sum."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import _abi  # noqa: E402

lib = _abi.load()
dev = torch.device("cuda:0")
lib.ospo_head_set_group_m(16)
cases = [
    (210, (4096, 1024, 16384), 16),          # 64 tiles, 1 wave on 64 clusters
    (210, (37 * 256, 512, 16384), 37),       # 74 tiles, 1 wave on 74 clusters
    (210, (8192, 1024, 16384), 16),          # 128 tiles: 74 + 54
    (210, (37 * 256, 1024, 16384), 37),      # 148 tiles: 2 full waves, each wave = 37 m x 2 n
    (210, (37 * 256, 2048, 16384), 37),      # 296 tiles: 4 full waves
    (210, (37 * 256, 4096, 16384), 37),      # 592 tiles: 8 full waves
    (210, (16384, 4096, 16384), 16),         # 1024 tiles, gm 16
    (210, (16384, 4096, 4096), 16),          # same tiles, short K
]
for variant, (M, N, K), gm in cases:
    lib.ospo_head_set_group_m(gm)
    majors = (variant // 10) % 10
    a_mn, b_mn = majors == 2, majors >= 1
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev)
    rc = lib.ospo_head_gemm_debug(variant, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), out.data_ptr(),
                                  out.stride(0), M, N, K, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ideal = (M * K + N * K) * 2 / 1e9
    print(f"WAVE variant={variant} M={M} N={N} K={K} gm={gm} rc={rc} once_GB={ideal:.3f}", flush=True)
    del A, B, out
