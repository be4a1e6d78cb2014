"""Timeline of the decode chain inside a CUDA graph (uses ospo_head_trace)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ospo_b200 import FusedGenHead, _abi, ops  # noqa: E402

import os

dev = torch.device("cuda:0")
H = E = int(os.environ.get("HE", "4096"))
ALT = int(os.environ.get("ALT", "1"))
DIRECT = int(os.environ.get("DIRECT", "0"))
N1 = int(os.environ.get("N1", "0"))   # chain prepare_gen_img_embeds behind the sampler (ospo_cfg_args.next_embeds)
V, P, steps = 16384, 16, 8


class Pm:
    n_embed, image_token_embed, image_token_size = H, E, V


torch.manual_seed(0)
head = FusedGenHead(Pm).to(dev).to(torch.bfloat16)
p = head._kernel_params()
alt = type(p)(p.w1.clone(), p.b1.clone(), p.w2.clone(), p.b2.clone())
PACK = int(os.environ.get("PACK", "1"))
pk = {id(p): (ops.pack_weight_impl(p.w1), ops.pack_weight_impl(p.w2)), id(alt): (ops.pack_weight_impl(alt.w1), ops.pack_weight_impl(alt.w2))}
h = torch.randn(steps, 2 * P, H, device=dev).to(torch.bfloat16)
u = torch.rand(steps, P, device=dev)
ids_out = torch.empty(steps, P, dtype=torch.int64, device=dev)
trace = torch.zeros(8, 160, 8, dtype=torch.int64, device=dev)
lib = _abi.load()


ne = None
if N1:
    from ospo_b200 import FusedGenImgEmbeds

    gen_embed = torch.nn.Embedding(V, 8).to(dev).to(torch.bfloat16)
    aligner = torch.nn.Module()
    aligner.layers = torch.nn.Sequential(torch.nn.Linear(8, H), torch.nn.GELU(), torch.nn.Linear(H, H)).to(dev).to(torch.bfloat16)
    fe = FusedGenImgEmbeds(gen_embed, aligner)
    ne = (*fe._params(), torch.empty(2 * P, H, dtype=torch.bfloat16, device=dev))


def run():
    for i in range(steps):
        w = p if (i & 1) == 0 or not ALT else alt
        if N1:
            ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i], ne, pk[id(w)])
        elif DIRECT:
            ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0, False, ids_out[i], None,
                                pk[id(w)] if PACK else None)
        else:
            ids, _ = ops.cfg_sample_impl(h[i], w.w1, w.b1, w.w2, w.b2, 5.0, 1.0, u[i], False, 0)
            ids_out[i].copy_(ids)


lib.ospo_head_trace(trace.data_ptr())  # before capture: the trace ids are launch parameters
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
trace.zero_()
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
lib.ospo_head_trace(None)
t = trace.cpu().numpy().astype("float64")
import os
merged = os.environ.get("OSPO_HEAD_DECODE_MERGED", "1") != "0"
names = {1: "phase1" if merged else "gemm1", 3: "prologue" if merged else "finalize", 2: "phase2" if merged else "gemm2", 4: "finish"}
slots_pro = ((0, "prologue_done"), (1, "partials_ready"), (2, "act_stored"), (3, "roles_done"), (4, "cta_exit"))
slots_merged = ((0, "entry"), (1, "A_prefetched"), (2, "wait_over"), (3, "w1_issued"), (4, "acc_ready"), (5, "parked"),
                (6, "published"), (7, "flag_passed"))
t0 = t[1][:, 0][t[1][:, 0] > 0].min()
for k in (1, 3, 2, 4):
    a = t[k]
    act = a[:, 0] > 0
    line = f"TRACE {names[k]:9s} ctas={int(act.sum()):3d}"
    for slot, nm in slots_merged if (merged and k == 1) else slots_pro if (merged and k == 3) else ((0, "entry"), (1, "prefetched"), (2, "wait_over"), (3, "prod_done"), (4, "acc_ready"), (5, "epi_done"), (6, "csync1"), (7, "finalized")):
        v = a[act, slot]
        v = v[v > 0]
        if v.size:
            line += f" | {nm} {((v.min() - t0) / 1e3):6.1f}..{((v.max() - t0) / 1e3):6.1f}"
    print(line)
for k, nm_k, slots in ((6, "aligner", slots_merged), (7, "align-pro", slots_pro)):
    a = t[k]
    act = a[:, 0] > 0
    if act.any():
        line = f"TRACE {nm_k:9s} ctas={int(act.sum()):3d}"
        for slot, nm in slots:
            v = a[act, slot]
            v = v[v > 0]
            if v.size:
                line += f" | {nm} {((v.min() - t0) / 1e3):6.1f}..{((v.max() - t0) / 1e3):6.1f}"
        print(line)
a = t[5]
act = a[:, 0] > 0
if act.any():
    base = t[2][act, 4]  # acc_ready of the same CTA
    line = "TRACE gemm2-epilogue (us after the CTA's acc_ready, median/max):"
    for slot, nm in enumerate(("tmem_loaded", "merged", "kmax", "bar1", "exp_stored", "seg_summed", "bar2")):
        v = (a[act, slot] - base) / 1e3
        line += f" | {nm} {float(sorted(v)[len(v)//2]):.2f}/{v.max():.2f}"
    print(line)
print("(us relative to the first GEMM1 CTA entry of the last traced step)")
