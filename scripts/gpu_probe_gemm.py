"""GPU probe for the tcgen05 GEMM engine: every (cta_group, operand-major, tile) variant is checked
against a torch fp32 product in its own subprocess (a trap or watchdog in one variant cannot take
the others down), then timed on a large shape.  Writes gpurun_out/gemm_probe.jsonl.

    python scripts/gpu_probe_gemm.py            # all variants
    python scripts/gpu_probe_gemm.py --one 100  # a single variant, in-process
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

VARIANTS = [100, 110, 120, 101, 102, 200, 210, 220]
SHAPES = [
    (128, 256, 64),
    (256, 256, 128),
    (128, 256, 512),
    (384, 512, 256),
    (296, 520, 200),     # ragged M / N / K tails
    (1024, 2048, 1024),
    (4096, 4096, 512),   # many tiles per CTA: exercises ring wrap-around and both accumulator stages
]
PERF_SHAPES = [(16384, 16384, 4096), (16384, 4096, 16384), (16384, 4096, 73728 // 4)]


def run_one(variant: int, perf: bool) -> list[dict]:
    import torch

    from ospo_b200 import _abi

    lib = _abi.load()
    dev = torch.device("cuda:0")
    majors = (variant // 10) % 10
    a_mn = majors == 2
    b_mn = majors >= 1
    out_rows = []
    g = torch.Generator(device="cpu").manual_seed(1234 + variant)

    def make(M, N, K):
        A = torch.randn(M, K, generator=g).to(torch.bfloat16)
        B = torch.randn(N, K, generator=g).to(torch.bfloat16)
        return A, B

    def call(A_dev, B_dev, out, M, N, K):
        lda = A_dev.stride(0)
        ldb = B_dev.stride(0)
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.ospo_head_gemm_debug(variant, A_dev.data_ptr(), lda, B_dev.data_ptr(), ldb, out.data_ptr(),
                                      out.stride(0), M, N, K, st)
        return rc

    for (M, N, K) in SHAPES:
        A, B = make(M, N, K)
        ref = (A.to(dev).float() @ B.to(dev).float().t())
        A_dev = (A.t().contiguous() if a_mn else A).to(dev)   # MN-major: stored [K, M]
        B_dev = (B.t().contiguous() if b_mn else B).to(dev)
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
        rec = {"variant": variant, "shape": [M, N, K], "kind": "check"}
        try:
            rc = call(A_dev, B_dev, out, M, N, K)
            torch.cuda.synchronize()
            rec["rc"] = rc
            if rc == 0:
                err = (out - ref).abs()
                bad = ~(err <= 1e-2 + 1e-3 * ref.abs())   # NaN counts as bad
                rec["max_abs_err"] = float(torch.nan_to_num(err, nan=1e30).max())
                rec["bad_frac"] = float(bad.float().mean())
                rec["nan_frac"] = float(torch.isnan(out).float().mean())
                if bad.any():
                    idx = bad.nonzero()[0].tolist()
                    rec["first_bad"] = idx
                    rec["first_bad_got_ref"] = [float(out[idx[0], idx[1]]), float(ref[idx[0], idx[1]])]
                    # coarse map: fraction bad per (row-block of 32, col-block of 32), first 8x8 blocks
                    mb, nb = min(8, (M + 31) // 32), min(8, (N + 31) // 32)
                    cm = []
                    for i in range(mb):
                        row = []
                        for j in range(nb):
                            blk = bad[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32]
                            row.append(round(float(blk.float().mean()), 2) if blk.numel() else -1)
                        cm.append(row)
                    rec["bad_map_32x32"] = cm
                rec["ok"] = bool(not bad.any())
            else:
                rec["ok"] = False
                rec["err"] = _abi.strerror(rc)
        except Exception as e:  # CUDA error (trap, illegal address, ...)
            rec["ok"] = False
            rec["exc"] = repr(e)[:400]
            rec["watchdog"] = _abi.watchdog_record()
            out_rows.append(rec)
            return out_rows
        out_rows.append(rec)
        if not rec["ok"]:
            return out_rows  # stop at the first failing shape

    if perf:
        for (M, N, K) in PERF_SHAPES:
            A_dev = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
            B_dev = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
            out = torch.empty((M, N), device=dev, dtype=torch.float32)
            for gm in (4, 8, 16):
                lib.ospo_head_set_group_m(gm)
                for _ in range(2):
                    call(A_dev, B_dev, out, M, N, K)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 5
                ev0.record()
                for _ in range(iters):
                    call(A_dev, B_dev, out, M, N, K)
                ev1.record()
                torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / iters
                out_rows.append({"variant": variant, "shape": [M, N, K], "kind": "perf", "group_m": gm, "ms": ms,
                                 "tflops": 2.0 * M * N * K / ms / 1e9})
            # cuBLAS for scale (library reference, not the product)
            Af = A_dev.t() if a_mn else A_dev
            Bf = B_dev.t() if b_mn else B_dev
            for _ in range(2):
                torch.matmul(Af, Bf.t())
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(5):
                torch.matmul(Af, Bf.t())
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 5
            out_rows.append({"variant": variant, "shape": [M, N, K], "kind": "cublas", "ms": ms,
                             "tflops": 2.0 * M * N * K / ms / 1e9})
    return out_rows


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", type=int, default=None)
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--variants", type=str, default=None)
    args = ap.parse_args()
    if args.one is not None:
        for r in run_one(args.one, args.perf):
            print("PROBE " + json.dumps(r), flush=True)
        return
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    variants = [int(v) for v in args.variants.split(",")] if args.variants else VARIANTS
    with open(out_dir / "gemm_probe.jsonl", "a") as f:
        for v in variants:
            t0 = time.time()
            cmd = [sys.executable, str(Path(__file__).resolve()), "--one", str(v)] + (["--perf"] if args.perf else [])
            try:
                p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
                lines = [l[6:] for l in p.stdout.splitlines() if l.startswith("PROBE ")]
                for l in lines:
                    f.write(l + "\n")
                    print(l)
                if p.returncode != 0:
                    msg = {"variant": v, "kind": "proc", "returncode": p.returncode, "stderr": p.stderr[-1500:]}
                    f.write(json.dumps(msg) + "\n")
                    print(json.dumps(msg))
            except subprocess.TimeoutExpired:
                msg = {"variant": v, "kind": "proc", "timeout": True}
                f.write(json.dumps(msg) + "\n")
                print(json.dumps(msg))
            f.flush()
            print(f"# variant {v} took {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
