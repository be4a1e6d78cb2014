"""bench.py -- OSPO image-token head on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one SimPO head forward+backward (BASELINE.json configs[1]: Janus-Pro-7B-shaped gen_head,
64 preference pairs x 576 image tokens per GPU, bf16, head trainable) over one batch of synthetic
hidden states.  N > 1 (torchrun, one rank per GPU) shards pairs across ranks -- 64 pairs per GPU, i.e.
configs[2]'s 512 pairs at N = 8 -- and all-reduces the flat fp32 head-weight gradient over NCCL
(weak scaling).  Rank 0 prints ONE JSON line.

  value      image-tokens/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        the same through FusedGenHead.simpo with HOST (pinned) inputs: every step copies its
             hidden states + labels host->device and reads the loss back, copies overlapped with compute
  roofline   dominant kernel (longest total time in the timed region), algorithmic flops / measured time
  cpu_baseline   the oracle (torch CPU restatement of the reference path) on a bounded sample
  cfg        secondary metric: BASELINE.json configs[3], 576 sequential CFG decode steps at P = 16

``--impl reference`` times the reference's CPU path (the oracle port) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H7B, E7B, V = 4096, 4096, 16384
T_IMG = 576
PAIRS_PER_GPU = 64
HP = dict(beta=10.0, gamma_beta_ratio=0.5, label_smoothing=0.0, sft_weight=0.0, loss_type="sigmoid")
METRIC = "simpo_head_fwd_bwd_image_tokens_per_s"
UNIT = "image-tokens/s"
WORKLOAD = ("configs[1]: Janus-Pro-7B-shaped gen_head (H=E=4096, V=16384) SimPO fwd+bwd, 64 synthetic pairs x 576 "
            "tokens per GPU (73728 rows), bf16, head trainable")


def flops_per_token(H, E, Vv):
    return 6 * (H * E + E * Vv)   # fwd 2(HE+EV) + bwd 4(HE+EV)


# -------------------------------------------------------------------------------------------------
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained"),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) -- used for cpu_baseline and for --impl reference
# -------------------------------------------------------------------------------------------------
def cpu_simpo_sample(pairs: int, reps: int):
    import torch

    from oracle import head_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    head = O.make_head(H7B, E7B, V, seed=1235)
    hc, hr, lc, lr = O.synthetic_simpo_batch(pairs, T_IMG, 1, H7B, V, seed=1236)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.simpo_step(head, hc, hr, lc, lr, backward=True, **HP)
        times.append(time.perf_counter() - t0)
    tokens = 2 * pairs * T_IMG
    return tokens, times, cores


def run_reference_arm(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = 4
    tokens, times, cores = cpu_simpo_sample(pairs, args.warmup + args.steps)
    timed = times[args.warmup:] or times
    ms = 1e3 * sum(timed) / len(timed)
    value = tokens / (ms / 1e3)
    sample = (f"{pairs} pairs x {T_IMG} tokens ({tokens} rows) of the 7B-shaped head, fp32 torch CPU oracle port of the "
              f"reference path, fwd+bwd, {cores} threads")
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours")
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU (default = configs[1])")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-cfg", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: library chatter (e.g. NCCL's version banner) goes to stderr meanwhile
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args, emit)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from ospo_b200 import FusedGenHead, _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (no CPU path for the product arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    peaks = load_peaks()
    B = args.pairs
    rows = 2 * B * T_IMG
    tokens_per_step_rank = rows
    L = 1   # one leading masked position so the label shift of train.py:385-386 is exercised

    # ---- synthetic inputs (seeded; weights identical on every rank, data differs per rank) ------
    torch.manual_seed(1235)

    class P:
        n_embed, image_token_embed, image_token_size = H7B, E7B, V

    head = FusedGenHead(P).to(dev).to(torch.bfloat16)        # default nn.Linear init under the seed
    gen = torch.Generator(device=dev).manual_seed(1236 + rank)
    hidden = torch.randn(2 * B, L + T_IMG, H7B, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ids = torch.randint(0, V, (2 * B, T_IMG), generator=gen, device=dev)
    labels = torch.cat([torch.full((2 * B, L), -100, dtype=torch.long, device=dev), ids], 1)
    span = (L - 1, L - 1 + T_IMG)

    def step(h, lab):
        head.zero_grad(set_to_none=True)
        hh = h.detach().requires_grad_(True)
        out = head.simpo(hh, lab, image_span=span, process_group=group, **HP)
        out.loss.backward()
        return out.loss, hh.grad

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    for _ in range(args.warmup):
        step(hidden, labels)
    sync_all()
    _abi.profile_enable(True)
    _abi.profile_read()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _abi.load().ospo_head_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        loss, _ = step(hidden, labels)
    ev1.record()
    sync_all()
    t_wall1 = time.time()
    launches = int(_abi.load().ospo_head_launch_count() - launches0)
    prof = _abi.profile_read()
    _abi.profile_enable(False)
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_step = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms_step], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t)
    value = world * tokens_per_step_rank / (ms_step / 1e3)
    loss_val = float(loss.detach())

    # ---- per-kernel table + roofline of the dominant kernel ---------------------------------------
    HE, EV = H7B * E7B, E7B * V
    kflops = {  # algorithmic flops per launch
        "gemm1_bias_gelu": 2 * rows * HE, "gemm2_logits_lse": 2 * rows * EV, "dact_gelu_bwd": 2 * rows * EV,
        "wgrad_w2": 2 * rows * EV, "wgrad_w1": 2 * rows * HE, "dgrad_x": 2 * rows * HE,
    }
    kernels = {}
    for name, (tot_ms, n) in prof.items():
        per = tot_ms / max(n, 1)
        ent = {"ms_per_launch": per, "launch_groups": n, "share_of_step": (tot_ms / args.steps) / ms_step}
        if name in kflops:
            ent["tflops"] = kflops[name] / per / 1e9
            ent["frac_of_peak"] = ent["tflops"] / peaks["tf_burst"]
        kernels[name] = ent
    gemm_names = [k for k in kernels if k in kflops]
    dominant = max(gemm_names, key=lambda k: kernels[k]["ms_per_launch"]) if gemm_names else None
    step_tflops = tokens_per_step_rank * flops_per_token(H7B, E7B, V) / (ms_step / 1e3) / 1e12
    roofline = None
    traffic, tensor_pct = None, None
    tf = ROOT / "profiles" / "r01_kernel_traffic.json"
    if dominant and tf.exists():
        ent = json.loads(tf.read_text()).get(dominant)
        if ent:
            traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
            tensor_pct = ent.get("tensor_pipe_active_pct")
    if dominant:
        roofline = {
            "bound": "tensor", "kernel": dominant, "achieved": kernels[dominant]["tflops"], "peak": peaks["tf_burst"],
            "unit": "TFLOP/s", "frac": kernels[dominant]["tflops"] / peaks["tf_burst"], "traffic": traffic,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of that kernel per launch, ncu capture of this "
                              "command (profiles/r01_kernel_traffic.json, r01_launches_ncu_v2.csv, "
                              "r01_dact_lockstep_ncu_summary.txt)" if traffic else None,
            "ncu_tensor_pipe_active_pct": tensor_pct,
            "peak_source": peaks["source"] + " bf16_tflops (burst); sustained " + str(peaks["tf_sustained"]),
            "step_achieved_tflops": step_tflops, "step_frac": step_tflops / peaks["tf_burst"],
            "note": "achieved = algorithmic flops of that launch / CUDA-event time around it, measured in the timed "
                    "region; step_* = 6(HE+EV) flop per image token over the whole fwd+bwd step",
        }

    # ---- end-to-end: host buffers through the public API -------------------------------------------
    e2e = None
    if not args.skip_e2e:
        NB = 2
        host_h = [torch.empty(hidden.shape, dtype=torch.bfloat16).pin_memory() for _ in range(NB)]
        host_l = [torch.empty(labels.shape, dtype=torch.long).pin_memory() for _ in range(NB)]
        for b in range(NB):
            host_h[b].copy_(hidden.cpu())
            host_l[b].copy_(labels.cpu())
        dev_h = [torch.empty_like(hidden) for _ in range(NB)]
        dev_l = [torch.empty_like(labels) for _ in range(NB)]
        host_loss = torch.empty(1, dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event() for _ in range(NB)]
        freed = [torch.cuda.Event() for _ in range(NB)]
        comp = torch.cuda.current_stream()

        def h2d(i):
            b = i % NB
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[b])
                dev_h[b].copy_(host_h[b], non_blocking=True)
                dev_l[b].copy_(host_l[b], non_blocking=True)
                ready[b].record(copy_stream)

        def run_e2e(n):
            for b in range(NB):
                freed[b].record(comp)
            h2d(0)
            for i in range(n):
                b = i % NB
                if i + 1 < n:
                    h2d(i + 1)                      # next step's inputs stream in while this step computes
                comp.wait_event(ready[b])
                l, _ = step(dev_h[b], dev_l[b])
                freed[b].record(comp)
                host_loss.copy_(l.detach().reshape(1), non_blocking=True)   # device -> host read of the result
            comp.synchronize()

        run_e2e(2)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        run_e2e(args.steps)
        e1.record()
        sync_all()
        wall_ms = (time.perf_counter() - tw0) * 1e3 / args.steps
        ms_e2e = max(e0.elapsed_time(e1) / args.steps, 0.0)
        ms_e2e = max(ms_e2e, wall_ms * 0.0)   # device time; wall clock reported beside it
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t)
        e2e = {"value": world * tokens_per_step_rank / (ms_e2e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": hidden.numel() * 2 + labels.numel() * 8, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e, "wall_ms_per_step": wall_ms,
               "api": "FusedGenHead.simpo(hidden, labels) + loss.backward(); pinned host inputs, double-buffered H2D"}
        del dev_h, dev_l, host_h, host_l

    # ---- secondary: the reference's default training configuration freezes gen_head (configs/step5.yaml:59-66):
    #      only dX leaves the head, 4 (HE + EV) flop per image token ---------------------------------
    frozen = None
    if rank == 0 and world == 1 and not args.skip_cfg:
        for prm in head.parameters():
            prm.requires_grad_(False)
        for _ in range(3):
            step(hidden, labels)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = max(3, min(10, args.steps))
        f0.record()
        for _ in range(nrep):
            step(hidden, labels)
        f1.record()
        torch.cuda.synchronize()
        ms_f = f0.elapsed_time(f1) / nrep
        tf_f = tokens_per_step_rank * (4 * (H7B * E7B + E7B * V)) / (ms_f / 1e3) / 1e12
        frozen = {"workload": "same batch, gen_head frozen (the reference's default, configs/step5.yaml:59-66): forward + dX only",
                  "ms_per_step": ms_f, "image_tokens_per_s": tokens_per_step_rank / (ms_f / 1e3), "tflops": tf_f,
                  "frac_of_peak": tf_f / peaks["tf_burst"]}
        for prm in head.parameters():
            prm.requires_grad_(True)

    # ---- secondary: CFG decode (configs[3]) ---------------------------------------------------------
    cfg = None
    if rank == 0 and not args.skip_cfg:
        cfg = bench_cfg(head, dev, peaks, with_cpu=(world == 1 and not args.skip_cpu))

    # ---- secondary: clip + AdamW on the flat buffers (next row N3) ------------------------------------
    opt_res = None
    if rank == 0 and not args.skip_cfg:
        try:
            opt_res = bench_clip_adamw(dev, peaks)
        except Exception as ex:  # a secondary measurement must not take the headline line down
            opt_res = {"error": repr(ex)[:200]}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        pairs = 4
        tokens, times, cores = cpu_simpo_sample(pairs, 2)
        best = min(times)
        cpu = {"value": tokens / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{pairs} pairs x {T_IMG} tokens ({tokens} rows) of the same 7B-shaped head, fp32 torch CPU "
                         f"oracle (port of the reference path), fwd+bwd, best of 2, {cores} threads"}

    if rank == 0:
        emit(({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": B, "global_pairs": B * world, "rows_per_gpu": rows,
                       "H": H7B, "E": E7B, "V": V, "parallelism": f"dp{world}: pairs batch-sharded, NCCL all-reduce of "
                       "the flat fp32 head gradient (83.9M elements)" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (604 MB hidden states + 2.4 GB bf16 logits spill per step)",
                       "cta_group": _abi.load().ospo_head_set_cta_group(0), "loss": loss_val},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
            "cpu_baseline": cpu, "frozen_head": frozen, "cfg": cfg, "clip_adamw": opt_res,
        }))
    if world > 1:
        dist.destroy_process_group()


def bench_clip_adamw(dev, peaks):
    """next row N3: one optimizer step for the 7B-shaped head (83.9 M fp32 elements): squared-norm pass (4 B/elem
    read) + fused clip/AdamW pass (g, p, m, v read; p, m, v + bf16 operand shadow written: 30 B/elem).  The five
    buffers total 1.6 GB, far beyond L2."""
    import torch

    from ospo_b200 import ops

    n = ops.flat_grad_numel(H7B, E7B, V)
    gen = torch.Generator(device=dev).manual_seed(7)
    g = torch.randn(n, generator=gen, device=dev) * 1e-3
    p = torch.randn(n, generator=gen, device=dev) * 0.02
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    shadow = torch.empty(V * E7B + E7B * H7B, dtype=torch.bfloat16, device=dev)

    def step(t):
        sq = ops.grad_sqnorm_impl(g)
        ops.adamw_step_impl(g, p, m, v, t, 4e-5, 0.9, 0.95, 1e-8, 0.0, 1.0, sq, shadow)

    for t in range(1, 4):
        step(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for t in range(4, 4 + reps):
        step(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = n * 34 - (n - shadow.numel()) * 2
    return {"workload": "clip_grad_norm_(1.0) + AdamW step on the flat fp32 buffers of the 7B-shaped head (83.9 M elements)",
            "ms_per_step": ms, "bytes_per_step": nbytes,
            "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm"]}}


def _decode_traffic():
    """dram bytes (read + write) of one decode-step launch from the committed ncu --set full capture, or None"""
    tf = ROOT / "profiles" / "r01_kernel_traffic.json"
    try:
        t = json.loads(tf.read_text())["decode_merged"]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def cpu_cfg_sample(steps: int):
    """the reference's decode step (gen_head + CFG merge + softmax + multinomial, image_generation.py:156-163) on the
    host cores: 7B-shaped head in bf16 like the generation path (utils/model.py:39), P = 16"""
    import torch

    from oracle import head_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = 16
    head = O.make_head(H7B, E7B, V, seed=1237).to(torch.bfloat16)
    g = torch.Generator().manual_seed(1238)
    h = torch.randn(steps + 1, 2 * P, H7B, generator=g).to(torch.bfloat16)
    times = []
    with torch.no_grad():
        for i in range(steps + 1):
            t0 = time.perf_counter()
            O.decode_step_reference(head, h[i], 5.0, 1.0, generator=g)
            times.append(time.perf_counter() - t0)
    return P, times[1:], cores


def bench_cfg(head, dev, peaks, with_cpu=False):
    """configs[3]: P = 16 pairs (32 CFG rows), cfg_weight 5, temperature 1, 576 sequential decode steps of
    gen_head + merge + sample.  The 576 steps are captured in one CUDA graph (the loop is launch-bound
    otherwise); weights are re-read from HBM every step because two weight copies alternate and each step
    also streams its own hidden-state slab."""
    import torch

    from ospo_b200 import cfg_merge_sample, ops

    P, steps = 16, T_IMG
    p = head._kernel_params()
    # a second copy of the weights so consecutive steps cannot hit the previous step's lines in L2
    alt = type(p)(p.w1.clone(), p.b1.clone(), p.w2.clone(), p.b2.clone())
    # the decode kernel's streaming layout of both weight copies (what FusedGenHead.cfg_sample keeps cached)
    packed = {id(p): (ops.pack_weight_impl(p.w1), ops.pack_weight_impl(p.w2)),
              id(alt): (ops.pack_weight_impl(alt.w1), ops.pack_weight_impl(alt.w2))}
    gen = torch.Generator(device=dev).manual_seed(1238)
    h = torch.randn(steps, 2 * P, H7B, generator=gen, device=dev).to(torch.bfloat16)
    u = torch.rand(steps, P, generator=gen, device=dev)
    ids_out = torch.empty(steps, P, dtype=torch.int64, device=dev)

    def run_steps():
        for i in range(steps):
            w = p if (i & 1) == 0 else alt
            # the ids land directly in the generated-token buffer (generated_tokens[:, i], image_generation.py:164)
            ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i], None,
                                packed[id(w)])

    run_steps()
    torch.cuda.synchronize()
    result = {}
    # eager (one launch group per step)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps()
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1)
    # CUDA graph of the whole 576-step loop
    graph_ms = None
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run_steps()
        torch.cuda.current_stream().wait_stream(s)
        gobj = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gobj):
            run_steps()
        gobj.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            gobj.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / 3
    except Exception as ex:  # graph capture is an optimisation, not a requirement
        result["graph_error"] = repr(ex)[:200]
    best_ms = graph_ms if graph_ms is not None else eager_ms
    step_bytes = 2 * (H7B * E7B + E7B * V) + 4 * (E7B + V) + 2 * 2 * P * H7B + 4 * P + 8 * P
    us_step = best_ms * 1e3 / steps
    result.update({
        "workload": "configs[3]: P=16 cond/uncond pairs, cfg_weight=5, temperature=1, 576 decode steps, 7B-shaped head",
        "tokens_per_s": steps * P / (best_ms / 1e3), "us_per_step": us_step,
        "eager_us_per_step": eager_ms * 1e3 / steps, "graph_us_per_step": None if graph_ms is None else graph_ms * 1e3 / steps,
        "roofline": {"bound": "hbm", "achieved": step_bytes / (us_step * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": step_bytes / (us_step * 1e-6) / 1e9 / peaks["hbm"], "bytes_per_step": step_bytes,
                     "traffic": _decode_traffic()},
    })
    if with_cpu:
        try:
            Pc, times, cores = cpu_cfg_sample(10)
            med = statistics.median(times)
            result["cpu_baseline"] = {
                "value": Pc / med, "unit": "tokens/s", "cores": cores, "kind": "port",
                "sample": f"median of {len(times)} decode steps (after 1 warm-up) of the same P=16, 7B-shaped head, bf16 torch "
                          f"CPU oracle (port of image_generation.py:156-163), {cores} threads", "us_per_step": med * 1e6}
        except Exception as ex:
            result["cpu_baseline"] = {"error": repr(ex)[:200]}
    # next row N1: the same loop with prepare_gen_img_embeds (gen_embed -> gen_aligner, +33.6 MB of weights) fused in
    try:
        from ospo_b200 import FusedGenImgEmbeds

        gen_embed = torch.nn.Embedding(V, 8).to(dev).to(torch.bfloat16)
        lin = torch.nn.Sequential(torch.nn.Linear(8, H7B), torch.nn.GELU(), torch.nn.Linear(H7B, H7B))
        aligner = torch.nn.Module()
        aligner.layers = lin.to(dev).to(torch.bfloat16)
        fused_embeds = FusedGenImgEmbeds(gen_embed, aligner)
        emb_out = torch.empty(2 * P, H7B, dtype=torch.bfloat16, device=dev)
        ne = (*fused_embeds._params(), emb_out)

        def run_steps_n1():
            for i in range(steps):
                w = p if (i & 1) == 0 else alt
                # image_generation.py:156-168 as one launch chain: decode kernel, finish (+ first aligner layer), D x D Linear
                ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i], ne,
                                    packed[id(w)])

        run_steps_n1()
        torch.cuda.synchronize()
        s2 = torch.cuda.Stream()
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            run_steps_n1()
        torch.cuda.current_stream().wait_stream(s2)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            run_steps_n1()
        g2.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            g2.replay()
        e1.record()
        torch.cuda.synchronize()
        us_n1 = e0.elapsed_time(e1) / 3 * 1e3 / steps
        bytes_n1 = step_bytes + 2 * H7B * H7B + 2 * 8 * H7B + 8 * H7B + 2 * 2 * P * H7B
        result["with_gen_img_embeds"] = {"us_per_step": us_n1, "bytes_per_step": bytes_n1,
                                         "achieved_gbs": bytes_n1 / (us_n1 * 1e-6) / 1e9,
                                         "frac_of_hbm_peak": bytes_n1 / (us_n1 * 1e-6) / 1e9 / peaks["hbm"]}
    except Exception as ex:
        result["with_gen_img_embeds"] = {"error": repr(ex)[:200]}
    # the 1B-shaped head (H = E = 2048, configs[0]'s shape): 75.6 MB of weights per step, four copies in rotation so
    # that 227 MB of other weights pass through the 126 MB L2 between two uses of a copy
    try:
        H1 = E1 = 2048
        g1 = torch.Generator(device=dev).manual_seed(1239)
        copies = []
        for _ in range(4):
            w1 = (torch.randn(E1, H1, generator=g1, device=dev) * H1 ** -0.5).to(torch.bfloat16)
            w2 = (torch.randn(V, E1, generator=g1, device=dev) * E1 ** -0.5).to(torch.bfloat16)
            b1 = torch.zeros(E1, dtype=torch.float32, device=dev)
            b2 = torch.zeros(V, dtype=torch.float32, device=dev)
            copies.append((w1, b1, w2, b2, (ops.pack_weight_impl(w1), ops.pack_weight_impl(w2))))
        h1 = torch.randn(steps, 2 * P, H1, generator=g1, device=dev).to(torch.bfloat16)

        def run_steps_1b():
            for i in range(steps):
                w1, b1, w2, b2, pk = copies[i & 3]
                ops.cfg_sample_impl(h1[i], w1, b1, w2, b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i], None, pk)

        run_steps_1b()
        torch.cuda.synchronize()
        s3 = torch.cuda.Stream()
        s3.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s3):
            run_steps_1b()
        torch.cuda.current_stream().wait_stream(s3)
        g3 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g3):
            run_steps_1b()
        g3.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            g3.replay()
        e1.record()
        torch.cuda.synchronize()
        us_1b = e0.elapsed_time(e1) / 3 * 1e3 / steps
        bytes_1b = 2 * (H1 * E1 + E1 * V) + 4 * (E1 + V) + 2 * 2 * P * H1 + 4 * P + 8 * P
        result["shape_1b"] = {"workload": "same loop, 1B-shaped head (H=E=2048)", "us_per_step": us_1b,
                              "tokens_per_s": P / (us_1b * 1e-6), "bytes_per_step": bytes_1b,
                              "achieved_gbs": bytes_1b / (us_1b * 1e-6) / 1e9,
                              "frac_of_hbm_peak": bytes_1b / (us_1b * 1e-6) / 1e9 / peaks["hbm"]}
        del copies, h1
    except Exception as ex:
        result["shape_1b"] = {"error": repr(ex)[:200]}
    # merge + sample alone on supplied logits, all 576 steps in one launch (SURVEY §8d secondary metric)
    lg = (torch.randn(steps, 2 * P, V, generator=gen, device=dev) * 3).to(torch.bfloat16)
    for _ in range(3):
        cfg_merge_sample(lg, 5.0, 1.0, uniforms=u)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        cfg_merge_sample(lg, 5.0, 1.0, uniforms=u)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = lg.numel() * 2 + steps * P * 12
    result["merge_sample_only"] = {"ms": ms, "tokens_per_s": steps * P / (ms / 1e3), "achieved_gbs": nbytes / ms / 1e6,
                                   "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm"], "bytes": nbytes}
    return result


if __name__ == "__main__":
    main()
