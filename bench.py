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
``--impl torch_gpu`` (context, not the anchor) runs the same step as plain PyTorch on the GPU -- what OSPO executes
today: bf16 Linear / GELU / Linear through cuBLAS, fp32 log_softmax + gather, autograd -- and reports its step time
and peak memory next to the fused path's.
``--global-pairs G`` fixes the TOTAL batch (strong scaling, configs[2]: 512 pairs over 2 / 4 / 8 GPUs).
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

    def start(self, wait_s: float = 5.0):
        """start sampling every 50 ms and wait until the first sample has arrived (nvidia-smi takes a few hundred ms
        to come up -- longer than a short timed region -- so the sampler is started before the warm-up steps)"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t_end = time.time() + wait_s
            while not self.rows and time.time() < t_end:
                time.sleep(0.02)
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
        rows = self.rows
        if t0 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t1 + 0.1]
            if not inside and rows:
                # timed region shorter than the sampling period: take the sample nearest to it
                mid = 0.5 * (t0 + t1)
                inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            rows = inside
        for ts, line in rows:
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
CPU_PAIRS = 8   # the 7B shape at 8 pairs (per-token figure comparable to configs[1])


def cpu_simpo_sample(pairs: int, reps: int, hidden: int = H7B):
    import torch

    from oracle import head_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    head = O.make_head(hidden, hidden, V, seed=1235)
    hc, hr, lc, lr = O.synthetic_simpo_batch(pairs, T_IMG, 1, hidden, V, seed=1236)
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
    pairs = CPU_PAIRS
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
# context arm: the same step as plain PyTorch on the GPU (what OSPO runs today on the same B200)
# -------------------------------------------------------------------------------------------------
def torch_head_simpo_step(params, hidden, labels, B):
    """janus/models/modeling_vlm.py:47-51 + ospo/wrapper/train.py:375-396, 317-342, 419 in stock PyTorch: bf16 Linear /
    GELU / Linear under autocast (cuBLAS), log_softmax promoted to fp32 by autocast, gather, masked mean, SimPO
    sigmoid loss, autograd backward.  Written out here (not imported from oracle/) because it is a GPU timing arm."""
    import torch
    import torch.nn.functional as F

    W1, b1, W2, b2 = params
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = F.linear(F.gelu(F.linear(hidden, W1, b1)), W2, b2)
        lab = labels[:, 1:].clone()
        lg = logits[:, :-1, :]
        mask = lab != -100
        lab[lab == -100] = 0
        ptl = torch.gather(lg.log_softmax(-1), dim=2, index=lab.unsqueeze(2)).squeeze(2)   # fp32 under autocast
        logps = (ptl * mask).sum(-1) / mask.sum(-1)
        z = (logps[:B] - logps[B:]) - HP["gamma_beta_ratio"]
        loss = (-F.logsigmoid(HP["beta"] * z)).mean()
    loss.backward()
    return loss.detach()


def run_torch_gpu_arm(args, emit):
    import torch

    from ospo_b200 import FusedGenHead

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl torch_gpu needs a GPU")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, L = args.pairs, 1
    rows = 2 * B * T_IMG
    torch.manual_seed(1235)

    class P:
        n_embed, image_token_embed, image_token_size = H7B, E7B, V

    head = FusedGenHead(P).to(dev).to(torch.bfloat16)
    gen = torch.Generator(device=dev).manual_seed(1236)
    hidden = torch.randn(2 * B, L + T_IMG, H7B, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ids = torch.randint(0, V, (2 * B, T_IMG), generator=gen, device=dev)
    labels = torch.cat([torch.full((2 * B, L), -100, dtype=torch.long, device=dev), ids], 1)
    params = [head.output_mlp_projector.weight, head.output_mlp_projector.bias, head.vision_head.weight,
              head.vision_head.bias]

    def torch_step():
        head.zero_grad(set_to_none=True)
        hh = hidden.detach().requires_grad_(True)
        return torch_head_simpo_step(params, hh, labels, B)

    def fused_step():
        head.zero_grad(set_to_none=True)
        hh = hidden.detach().requires_grad_(True)
        out = head.simpo(hh, labels, image_span=(L - 1, L - 1 + T_IMG), **HP)
        out.loss.backward()
        return out.loss.detach()

    def measure(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        resident = torch.cuda.memory_allocated(dev)
        torch.cuda.reset_peak_memory_stats(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, torch.cuda.max_memory_allocated(dev), resident, float(loss)

    clocks = ClockSampler(0)
    clocks.start()
    torch_step()
    torch.cuda.synchronize()
    t0 = time.time()
    ms_t, peak_t, res_t, loss_t = measure(torch_step)
    t1 = time.time()
    clk = clocks.stop(t0, t1)
    head.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    ms_f, peak_f, res_f, loss_f = measure(fused_step)
    value = rows / (ms_t / 1e3)
    emit({
        "impl": "torch_gpu", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": B, "rows_per_gpu": rows,
                   "what": "stock PyTorch on the same GPU: bf16 autocast Linear/GELU/Linear (cuBLAS), fp32 log_softmax + "
                           "gather, SimPO sigmoid loss, autograd; head trainable, inputs resident in HBM", "loss": loss_t},
        "clocks": clk, "gpu_launches": 0,
        "memory": {"peak_bytes": peak_t, "resident_before_step_bytes": res_t, "step_transient_bytes": peak_t - res_t},
        "fused_same_process": {"ms_per_step": ms_f, "value": rows / (ms_f / 1e3), "loss": loss_f,
                               "memory": {"peak_bytes": peak_f, "resident_before_step_bytes": res_f,
                                          "step_transient_bytes": peak_f - res_f},
                               "speedup_vs_torch_gpu": ms_t / ms_f,
                               "transient_memory_ratio_torch_over_fused": (peak_t - res_t) / max(1, peak_f - res_f)},
    })



# -------------------------------------------------------------------------------------------------
# configs[4]: end-to-end SimPO training step, 7B-shaped Llama backbone + the head (fused vs PyTorch), N GPUs
# -------------------------------------------------------------------------------------------------
def run_config5(args, emit):
    """BASELINE.json configs[4]: random-init Janus-Pro-7B-shaped language model (HF LlamaModel, 30 layers, hidden 4096,
    32 heads, intermediate 11008, vocab 102400; activation checkpointing, trainable, bf16) + gen_head frozen as in
    configs/step5.yaml:59-66, 16 pairs per GPU (step5.yaml:23), L = 24 text + 576 image positions.  One process per
    GPU; the backbone gradients are averaged by DDP (ospo/utils/train.py:26-28).  Two arms on the same weights and
    batch: (a) the reference formulation -- PyTorch vision_head on every position, fp32 log_softmax, gather
    (ospo/wrapper/train.py:345-445) -- and (b) ``patch_train_wrapper`` (fused head on the 576 useful rows).  Reports
    pairs/s of the whole job, the head's own time inside the step (CUDA events around the head forward / backward on
    the step's hidden states), and the peak memory of each arm."""
    import types

    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from transformers import LlamaConfig, LlamaModel

    from ospo_b200 import FusedGenHead, _abi, patch_train_wrapper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --config5 needs a B200")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    PAIRS = 16 if args.pairs == PAIRS_PER_GPU else args.pairs
    LAYERS = int(os.environ.get("OSPO_BENCH_LAYERS", "30"))
    L, T, H = 24, T_IMG, H7B
    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=H, intermediate_size=11008, num_hidden_layers=LAYERS, num_attention_heads=32,
                      num_key_value_heads=32, vocab_size=102400, max_position_embeddings=16384)
    cfg.output_hidden_states = True                       # ospo/wrapper/train.py:50
    with torch.device(dev):
        backbone = LlamaModel(cfg).to(torch.bfloat16)
    backbone.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    backbone.train()
    backbone.embed_tokens.weight.requires_grad_(False)    # the step feeds inputs_embeds (train.py:352): never used
    net = backbone
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(backbone, device_ids=[local_rank], gradient_as_bucket_view=True)

    class P:
        n_embed, image_token_embed, image_token_size = H, H, V

    torch.manual_seed(5)
    ref_head = torch.nn.Module()
    ref_head.output_mlp_projector = torch.nn.Linear(H, H)
    ref_head.vision_activation = torch.nn.GELU()
    ref_head.vision_head = torch.nn.Linear(H, V)
    ref_head.forward = lambda x: ref_head.vision_head(ref_head.vision_activation(ref_head.output_mlp_projector(x)))
    ref_head = ref_head.to(dev).to(torch.bfloat16)
    for prm in ref_head.parameters():
        prm.requires_grad_(False)                          # configs/step5.yaml:59-66
    g = torch.Generator().manual_seed(1 + rank)
    emb_c = (torch.randn(PAIRS, L + T, H, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    emb_r = (torch.randn(PAIRS, L + T, H, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    pad = torch.full((PAIRS, L), -100, dtype=torch.long)
    lab_c = torch.cat([pad, torch.randint(0, V, (PAIRS, T), generator=g)], 1).to(dev)
    lab_r = torch.cat([pad, torch.randint(0, V, (PAIRS, T), generator=g)], 1).to(dev)
    batch = {"chosen_inputs_embeds": emb_c, "chosen_labels": lab_c, "rejected_inputs_embeds": emb_r,
             "rejected_labels": lab_r}
    labels_all = torch.cat([lab_c, lab_r], 0)
    head_ev = {"fwd": [], "bwd": []}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def reference_head_loss(hidden):
        # ospo/wrapper/train.py:357 (head on EVERY position), :375-396, :317-342, :419
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = ref_head.forward(hidden)
            lab = labels_all[:, 1:].clone()
            lg = logits[:, :-1, :]
            mask = lab != -100
            lab[lab == -100] = 0
            ptl = torch.gather(lg.log_softmax(-1), dim=2, index=lab.unsqueeze(2)).squeeze(2)
            logps = (ptl * mask).sum(-1) / mask.sum(-1)
            z = (logps[:PAIRS] - logps[PAIRS:]) - HP["gamma_beta_ratio"]
            return (-F.logsigmoid(HP["beta"] * z)).mean()

    model = torch.nn.Module()
    model.language_model = torch.nn.Module()
    model.language_model.model = net
    model.gen_head = ref_head
    w = types.SimpleNamespace(model=model, label_pad_token_id=-100, logged={}, **HP)
    w.log = lambda name, val, **kw: w.logged.__setitem__(name, val)
    w.log_dict = lambda d, **kw: w.logged.update(d)
    w.concatenated_inputs = lambda batch: {
        "concatenated_inputs_embeds": torch.cat([batch["chosen_inputs_embeds"], batch["rejected_inputs_embeds"]], 0),
        "concatenated_labels": torch.cat([batch["chosen_labels"], batch["rejected_labels"]], 0)}

    def timed_head(loss_fn, hidden):
        """head forward + backward on the step's own hidden states, bracketed by events (detached leaf: the backbone's
        backward is run afterwards from the gradient the head returned)"""
        leaf = hidden.detach().requires_grad_(True)
        e0, e1, e2 = ev(), ev(), ev()
        step_peak = torch.cuda.max_memory_allocated(dev)
        before = torch.cuda.memory_allocated(dev)
        torch.cuda.reset_peak_memory_stats(dev)
        e0.record()
        loss = loss_fn(leaf)
        e1.record()
        loss.backward()
        e2.record()
        head_ev["fwd"].append((e0, e1))
        head_ev["bwd"].append((e1, e2))
        head_ev["mem"] = max(head_ev.get("mem", 0), torch.cuda.max_memory_allocated(dev) - before)
        head_ev["peak"] = max(head_ev.get("peak", 0), step_peak, torch.cuda.max_memory_allocated(dev))
        torch.cuda.reset_peak_memory_stats(dev)
        hidden.backward(leaf.grad)
        head_ev["peak"] = max(head_ev["peak"], torch.cuda.max_memory_allocated(dev))
        return loss.detach()

    def reference_step():
        hidden = net(inputs_embeds=torch.cat([emb_c, emb_r], 0), use_cache=False).hidden_states[-1]
        return timed_head(reference_head_loss, hidden)

    def fused_step():
        hidden = net(inputs_embeds=torch.cat([emb_c, emb_r], 0), use_cache=False).hidden_states[-1]
        return timed_head(lambda leaf: model.gen_head.simpo(leaf, labels_all, image_span=(L - 1, L - 1 + T),
                                                            **HP).loss, hidden)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def measure(fn):
        backbone.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        for _ in range(max(1, min(args.warmup, 2))):
            backbone.zero_grad(set_to_none=True)
            fn()
        backbone.zero_grad(set_to_none=True)
        head_ev["fwd"].clear()
        head_ev["bwd"].clear()
        head_ev["mem"] = head_ev["peak"] = 0
        sync_all()
        torch.cuda.reset_peak_memory_stats(dev)
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(args.steps):
            backbone.zero_grad(set_to_none=True)
            loss = fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1) / args.steps
        hf = sum(a.elapsed_time(b) for a, b in head_ev["fwd"]) / args.steps
        hb = sum(a.elapsed_time(b) for a, b in head_ev["bwd"]) / args.steps
        peak, head_mem = head_ev["peak"], head_ev["mem"]
        if world > 1:
            t = torch.tensor([ms, hf, hb, float(peak), float(head_mem)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, hf, hb, peak, head_mem = (float(v) for v in t)
        gnorm = backbone.layers[0].self_attn.q_proj.weight.grad.detach().float().clone()
        return {"ms_per_step": ms, "head_fwd_ms": hf, "head_bwd_ms": hb, "head_ms": hf + hb, "peak_bytes": int(peak),
                "head_transient_bytes": int(head_mem), "loss": float(loss)}, gnorm

    args.steps = max(1, min(args.steps, 5))
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t0 = time.time()
    ref_res, g_ref = measure(reference_step)
    patch_train_wrapper(w, image_span=(L - 1, L - 1 + T))
    assert isinstance(model.gen_head, FusedGenHead)
    launches0 = _abi.load().ospo_head_launch_count()
    fused_res, g_fused = measure(fused_step)
    launches = int(_abi.load().ospo_head_launch_count() - launches0)
    t1 = time.time()
    clk = clocks.stop(t0, t1) if rank == 0 else None
    grad_rel = float((g_fused - g_ref).norm() / g_ref.norm().clamp_min(1e-20))
    if rank == 0:
        for r_ in (ref_res, fused_res):
            r_["pairs_per_s"] = world * PAIRS / (r_["ms_per_step"] / 1e3)
            r_["head_share_of_step"] = r_["head_ms"] / r_["ms_per_step"]
        emit({
            "metric": "simpo_train_step_pairs_per_s", "value": fused_res["pairs_per_s"], "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": fused_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"configs[4]: random-init Janus-Pro-7B-shaped Llama backbone ({LAYERS} layers, hidden 4096, "
                                   f"activation checkpointing, trainable) + gen_head frozen, {PAIRS} pairs x ({L} text + 576 "
                                   f"image) positions per GPU, {world} GPU(s), DDP over the backbone gradients",
                       "pairs_per_gpu": PAIRS, "layers": LAYERS},
            "clocks": clk, "gpu_launches": launches, "fused_head": fused_res, "pytorch_head": ref_res,
            "head_speedup": ref_res["head_ms"] / fused_res["head_ms"],
            "head_transient_memory_ratio_pytorch_over_fused": ref_res["head_transient_bytes"] / max(1, fused_res["head_transient_bytes"]),
            "peak_memory_saved_bytes": ref_res["peak_bytes"] - fused_res["peak_bytes"],
            "memory_note": "peak_bytes is the step's maximum (weights + all backbone gradients at the end of the backward: "
                           "the head is not live then); head_transient_bytes is what the head's forward + backward "
                           "allocates on top of what is live when it starts",
            "parity": {"loss_fused": fused_res["loss"], "loss_pytorch_head": ref_res["loss"],
                       "first_layer_q_proj_grad_rel_err": grad_rel},
        })
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--global-pairs", type=int, default=0,
                    help="total pairs over all GPUs (strong scaling; configs[2] = 512); overrides --pairs")
    ap.add_argument("--config5", action="store_true",
                    help="configs[4]: Janus-Pro-7B-shaped random-init Llama backbone + the fused head, 16 pairs per GPU")
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
    if args.impl == "torch_gpu":
        run_torch_gpu_arm(args, emit)
        return
    if args.config5:
        run_config5(args, emit)
        return

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
    strong = args.global_pairs > 0
    if strong and args.global_pairs % world:
        raise SystemExit(f"--global-pairs {args.global_pairs} does not shard over {world} GPUs")
    B = args.global_pairs // world if strong else args.pairs
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
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step(hidden, labels)
    sync_all()
    _abi.profile_enable(True)
    _abi.profile_read()
    launches0 = _abi.load().ospo_head_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    mem_resident = torch.cuda.memory_allocated(dev)       # weights + inputs + the reusable flat gradient buffer
    torch.cuda.reset_peak_memory_stats(dev)
    t_wall0 = time.time()
    ev0.record()
    marks = []   # one event per step: the spread of the K steps is reported beside their total (step_ms)
    for _ in range(args.steps):
        loss, _ = step(hidden, labels)
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    ev1.record()
    sync_all()
    t_wall1 = time.time()
    per_step = sorted(a.elapsed_time(b) for a, b in zip([ev0] + marks[:-1], marks))
    step_ms = {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]}
    launches = int(_abi.load().ospo_head_launch_count() - launches0)
    mem_peak = torch.cuda.max_memory_allocated(dev)
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

    # ---- N > 1: value check of the gradient exchange (every rank's exchanged buffer is bit-identical and equals the
    #      mean of the pre-exchange local ones) -------------------------------------------------------------------
    dp = None
    if world > 1:
        from ospo_b200 import dist as D

        head.zero_grad(set_to_none=True)
        hh = hidden.detach().requires_grad_(True)
        head.simpo(hh, labels, image_span=span, process_group=None, **HP).loss.backward()
        flat_local = head._flat.clone()
        step(hidden, labels)
        dp = D.dp_check(flat_local, head._flat, group)
        del flat_local
        sync_all()
        # how much of the N-GPU step is waiting for the slowest board: the same step WITHOUT any exchange, timed on
        # every rank separately (a synchronous data-parallel step cannot be faster than its slowest rank's compute)
        def local_step():
            head.zero_grad(set_to_none=True)
            hh2 = hidden.detach().requires_grad_(True)
            head.simpo(hh2, labels, image_span=span, process_group=None, **HP).loss.backward()

        for _ in range(3):
            local_step()
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(5):
            local_step()
        l1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([l0.elapsed_time(l1) / 5], device=dev)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        local_ms = [float(t) for t in every]
        dp["local_step_ms_per_rank_no_exchange"] = local_ms
        dp["slowest_rank_local_ms"] = max(local_ms)
        dp["exchange_and_sync_overhead_ms"] = ms_step - max(local_ms)
        sync_all()

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
    tf = ROOT / "profiles" / "kernel_traffic.json"
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
                              "command (profiles/kernel_traffic.json, generated by scripts/make_kernel_traffic.py from "
                              "the launch list named inside it)" if traffic else None,
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
        ms_e2e = e0.elapsed_time(e1) / args.steps      # device time; the wall clock is reported beside it
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
        pairs = CPU_PAIRS
        tokens, times, cores = cpu_simpo_sample(pairs, 2)
        best = min(times)
        cpu = {"value": tokens / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{pairs} pairs x {T_IMG} tokens ({tokens} rows) of the same 7B-shaped head, fp32 torch CPU "
                         f"oracle (port of the reference path), fwd+bwd, best of 2, {cores} threads"}
        # BASELINE.json configs[0] exactly: 1B-shaped head (hidden 2048), 8 pairs x 576 tokens, fp32 on CPU
        tok1, t1, _ = cpu_simpo_sample(8, 4, hidden=2048)
        cpu["config1"] = {"workload": "configs[0]: 1B-shaped head (H=E=2048), 8 pairs x 576 tokens, fp32, fwd+bwd",
                          "value": tok1 / min(t1[1:]), "unit": UNIT, "s_per_step_best_of_3": min(t1[1:]),
                          "s_per_step_median_of_3": statistics.median(t1[1:])}

    if rank == 0:
        emit(({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD if not strong else
                       (f"configs[2]: Janus-Pro-7B-shaped SimPO head, {B * world} pairs batch-sharded over {world} GPU(s) "
                        f"({B} pairs x 576 tokens per GPU), bf16, head trainable, NCCL all-reduce of the head-weight "
                        "gradients"),
                       "pairs_per_gpu": B, "global_pairs": B * world, "rows_per_gpu": rows,
                       "H": H7B, "E": E7B, "V": V,
                       "parallelism": (f"dp{world}: pairs batch-sharded; head-weight gradients (83.9M fp32) exchanged "
                                       + ("over NVLink peer memory: reduce-scatter fused into the weight-gradient GEMM "
                                          "epilogues, owners sum in rank order and multicast their shards"
                                          if head._peer_exchange(group) is not None else "by NCCL all-reduce"))
                       if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (604 MB hidden states + 2.4 GB bf16 logits spill per step)",
                       "cta_group": _abi.load().ospo_head_set_cta_group(0), "loss": loss_val},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
            "step_ms": step_ms, "cpu_baseline": cpu, "frozen_head": frozen, "cfg": cfg, "clip_adamw": opt_res, "dp_check": dp,
            "memory": {"peak_bytes": mem_peak, "resident_before_step_bytes": mem_resident,
                       "step_transient_bytes": mem_peak - mem_resident,
                       "note": "torch.cuda.max_memory_allocated over the timed steps; resident = weights, inputs, "
                               "reusable flat gradient buffer; the reference path materialises fp32 logits + fp32 "
                               "log-softmax of [rows, V] (4.8 GB each at this size) plus autograd copies"},
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
    tf = ROOT / "profiles" / "kernel_traffic.json"
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
        # LATENCY of a truly sequential loop (image_generation.py:149-171: step i+1's hidden state depends on the id
        # step i sampled).  With D == H the aligner output [2P, D] of step i IS the hidden state of step i+1 (it stands
        # in for the backbone), so no kernel of step i+1 can consume anything before step i has produced it; the figure
        # above (independent hidden states) is the pipelined cadence, this one the dependent-step latency.
        emb_pp = [torch.empty(2 * P, H7B, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        ne_pp = [(*fused_embeds._params(), e) for e in emb_pp]

        def run_steps_dep():
            cur = h[0]
            for i in range(steps):
                w = p if (i & 1) == 0 else alt
                ops.cfg_sample_impl(cur, w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i],
                                    ne_pp[i & 1], packed[id(w)])
                cur = emb_pp[i & 1]

        run_steps_dep()
        torch.cuda.synchronize()
        s4 = torch.cuda.Stream()
        s4.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s4):
            run_steps_dep()
        torch.cuda.current_stream().wait_stream(s4)
        g4 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g4):
            run_steps_dep()
        g4.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            g4.replay()
        e1.record()
        torch.cuda.synchronize()
        us_dep = e0.elapsed_time(e1) / 3 * 1e3 / steps
        result["dependent_chain"] = {
            "workload": "576 steps, each step's hidden state = the previous step's sampled-id embeddings "
                        "(head -> merge+sample -> gen_embed -> gen_aligner -> next head): latency, not cadence",
            "us_per_step_dependent": us_dep, "us_per_step_pipelined": us_n1, "bytes_per_step": bytes_n1,
            "achieved_gbs": bytes_n1 / (us_dep * 1e-6) / 1e9,
            "frac_of_hbm_peak": bytes_n1 / (us_dep * 1e-6) / 1e9 / peaks["hbm"]}
        # the same dependent loop with the aligner memoised over the 16384 codes (FusedGenImgEmbeds.build_table): the
        # finish kernel copies the pair's two rows of the [16384, D] table, no aligner weight is streamed and no
        # further kernel runs.  Algorithmic bytes: the head's step + 2P table rows read + 2P rows written.
        table = fused_embeds.build_table()
        ne_tab = [(*fused_embeds._params(), e, table) for e in emb_pp]

        def run_steps_dep_table():
            cur = h[0]
            for i in range(steps):
                w = p if (i & 1) == 0 else alt
                ops.cfg_sample_impl(cur, w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i],
                                    ne_tab[i & 1], packed[id(w)])
                cur = emb_pp[i & 1]

        run_steps_dep_table()
        torch.cuda.synchronize()
        s5 = torch.cuda.Stream()
        s5.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s5):
            run_steps_dep_table()
        torch.cuda.current_stream().wait_stream(s5)
        g5 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g5):
            run_steps_dep_table()
        g5.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            g5.replay()
        e1.record()
        torch.cuda.synchronize()
        us_tab = e0.elapsed_time(e1) / 3 * 1e3 / steps
        bytes_tab = step_bytes + 2 * (2 * P * H7B) * 2
        result["dependent_chain_embed_table"] = {
            "workload": "the dependent loop with gen_aligner(gen_embed(id)) memoised per code (134 MB table): head -> "
                        "merge+sample -> two table rows per pair -> next head",
            "us_per_step_dependent": us_tab, "bytes_per_step": bytes_tab, "tokens_per_s": P / (us_tab * 1e-6),
            "achieved_gbs": bytes_tab / (us_tab * 1e-6) / 1e9,
            "frac_of_hbm_peak": bytes_tab / (us_tab * 1e-6) / 1e9 / peaks["hbm"],
            "speedup_vs_streamed_aligner": us_dep / us_tab}
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
