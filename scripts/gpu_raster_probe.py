"""DRAM traffic / time of the long-K backward GEMM shapes versus rasterisation group size (run under ncu with
--metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import _abi  # noqa: E402

lib = _abi.load()
dev = torch.device("cuda:0")
shapes = {210: (73728, 4096, 16384), 220: (16384, 4096, 73728), 200: (73728, 16384, 4096)}
gms = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 4, 8, 16, 32]
for variant, (M, N, K) in shapes.items():
    majors = (variant // 10) % 10
    a_mn, b_mn = majors == 2, majors >= 1
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev)
    for gm in gms:
        lib.ospo_head_set_group_m(gm)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = lib.ospo_head_gemm_debug(variant, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), out.data_ptr(),
                                      out.stride(0), M, N, K, torch.cuda.current_stream().cuda_stream)
        ev1.record()
        torch.cuda.synchronize()
        print(f"RASTER variant={variant} gm={gm} rc={rc} ms={ev0.elapsed_time(ev1):.3f}", flush=True)
    del A, B, out
