"""profiles/kernel_traffic.json from an ncu launch list (CSV with gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum and, optionally, the tensor-pipe and L2-hit metrics) of `bench.py --steps 2 --warmup 3
--skip-cpu --skip-e2e --skip-cfg`.  bench.py reads the file for `roofline.traffic`; the source CSV and the commit it
was captured at are recorded inside, so a stale file is recognisable.

    python scripts/make_kernel_traffic.py profiles/<launch list>.csv [decode launch list .csv]"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
TENSOR = "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    by = collections.OrderedDict()
    for r in rows:
        d = by.setdefault(r[iid], {"name": r[ik]})
        d[r[im]] = float(r[iv].replace(",", ""))
    return list(by.values())


def classify(d):
    n, us = d["name"], d.get("gpu__time_duration.sum", 0) / 1e3
    if "EpiBiasGelu" in n:
        return "gemm1_bias_gelu"
    if "EpiLogitsExp" in n:
        return "gemm2_logits_lse" if us > 1000 else "gemm2_repair_pass_empty"
    if "EpiDactScale" in n:
        return "dact_gelu_bwd"
    if "EpiRedAdd" in n:
        return "wgrad_w1"  # split-K form (two half-length work items per tile added into the zeroed output)
    if "EpiStore<float" in n:
        return "wgrad_w2" if us > 4000 else "wgrad_w1"
    if "EpiStore<__nv_bfloat16" in n and "gemm_kernel" in n:
        return "dgrad_x"
    if "colsum_partial_kernel" in n:
        return "colsum_db2" if d.get("dram__bytes_read.sum", 0) > 1.5e9 else "colsum_db1"
    if "decode_merged_kernel" in n:
        return "decode_merged"
    for k in ("lse_finalize_kernel", "target_fixup_kernel", "seq_reduce_kernel", "simpo_scalar_kernel",
              "row_weight_kernel", "colsum_final_kernel", "cfg_finish_kernel"):
        if k in n:
            return k
    return None


def main():
    src = [Path(p) for p in sys.argv[1:]]
    acc = collections.defaultdict(list)
    for p in src:
        for d in launches(p):
            k = classify(d)
            if k and d.get("gpu__time_duration.sum", 0) == d.get("gpu__time_duration.sum", 0):  # ncu prints nan for the
                acc[k].append(d)                                                                # cooperative decode kernel
    out = {"_source": [str(p.relative_to(ROOT)) if p.is_absolute() else str(p) for p in src],
           "_commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True,
                                     text=True).stdout.strip(),
           "_note": "per launch, averaged over the launches of the capture; ncu figures are serialised and cold-cache"}
    for k, ds in acc.items():
        n = len(ds)
        ent = {"launches": n,
               "dram_bytes_read": sum(d.get("dram__bytes_read.sum", 0) for d in ds) / n,
               "dram_bytes_write": sum(d.get("dram__bytes_write.sum", 0) for d in ds) / n,
               "time_us_under_ncu": sum(d.get("gpu__time_duration.sum", 0) for d in ds) / n / 1e3}
        if any(TENSOR in d for d in ds):
            ent["tensor_pipe_active_pct"] = sum(d.get(TENSOR, 0) for d in ds) / n
        if any("lts__t_sector_hit_rate.pct" in d for d in ds):
            ent["l2_hit_rate_pct"] = sum(d.get("lts__t_sector_hit_rate.pct", 0) for d in ds) / n
        out[k] = ent
    if "decode_merged" not in out:
        r1 = json.loads((ROOT / "profiles" / "r01_kernel_traffic.json").read_text()).get("decode_merged")
        if r1:
            out["decode_merged"] = {
                "launches": 1, "dram_bytes_read": r1["dram_bytes_read"], "dram_bytes_write": r1["dram_bytes_write"],
                "source": "carried over from profiles/r01_kernel_traffic.json (ncu --set full -k regex:decode_merged python "
                          "scripts/gpu_cfg_one.py, profiles/r01_decode_merged_ncu_summary.txt): the kernel's weight stream "
                          "is unchanged since; the launch list has no valid decode launch (a cooperative launch under "
                          "the metrics pass reports nan)"}
    (ROOT / "profiles" / "kernel_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps({k: (round(v["dram_bytes_read"] / 1e9, 2), round(v["dram_bytes_write"] / 1e9, 2))
                      for k, v in out.items() if not k.startswith("_")}))


if __name__ == "__main__":
    main()
