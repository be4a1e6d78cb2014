"""bench.py contract checks that need no GPU: the reference arm's JSON line, rank handling under torchrun, the
refusal of the product arm to run without a B200, and the keys of the committed end-of-round bench line."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                 "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline")


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in CONTRACT_KEYS:
        assert k in d, k
    assert d["metric"] == "simpo_head_fwd_bwd_image_tokens_per_s" and d["unit"] == "image-tokens/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = rows of the sample / time of one step
    assert d["value"] == pytest.approx(8 * 2 * 576 / (d["ms_per_step"] / 1e3), rel=1e-9)


def test_reference_arm_runs_on_rank_zero_only():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "3", "--skip-cfg", "--skip-cpu"])
    assert r.returncode != 0 and "needs a B200" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_committed_bench_line_carries_the_contract():
    line = (ROOT / "profiles" / "r02_bench_final6_n1.json").read_text().strip().splitlines()[-1]
    d = json.loads(line)
    for k in CONTRACT_KEYS + ("clocks", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert rf["frac"] == pytest.approx(rf["achieved"] / rf["peak"], rel=1e-9)
    assert d["value"] == pytest.approx(2 * 64 * 576 / (d["ms_per_step"] / 1e3), rel=1e-9)
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] != d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]
    dec = d["cfg"]["roofline"]
    assert dec["bound"] == "hbm" and dec["frac"] == pytest.approx(dec["achieved"] / dec["peak"], rel=1e-9)
    # round 2: no softmax-minus-onehot producer pass, peak memory reported, decode latency beside the cadence
    assert "dlogits_producer" not in d["kernels"] and d["memory"]["peak_bytes"] > 0
    dep = d["cfg"]["dependent_chain"]
    assert dep["us_per_step_dependent"] > d["cfg"]["us_per_step"]
    assert d["cfg"]["dependent_chain_embed_table"]["us_per_step_dependent"] < dep["us_per_step_dependent"]
    assert d["cpu_baseline"]["config1"]["value"] > 0
    # the K timed steps one by one: their total is the metric, their spread shows whether a step was an outlier
    sm = d["step_ms"]
    assert sm["min"] <= sm["median"] <= sm["max"] and sm["max"] < 1.1 * sm["min"]
    assert sm["min"] <= d["ms_per_step"] <= sm["max"]
    # north_star's literal "CFG merge+sample at >= 70 % of HBM bandwidth": the fused decode step and the stand-alone
    # sampler on supplied logits
    assert dec["frac"] >= 0.70 and d["cfg"]["merge_sample_only"]["frac_of_hbm_peak"] >= 0.70
    assert d["roofline"]["step_frac"] >= 0.60


def test_committed_multi_gpu_lines_carry_a_passed_dp_check():
    for name in ("r02_bench_n2.json", "r02_bench_n2_final3.json", "r02_bench_n8_nccl.json", "r02_bench_n8_p2p2.json",
                 "r02_bench_n8_final.json"):
        d = json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])
        assert d["n_gpus"] > 1 and d["dp_check"]["status"] == "ok" and d["dp_check"]["ranks_bit_identical"] is True
        assert d["dp_check"]["world"] == d["n_gpus"]
