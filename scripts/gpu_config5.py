"""BASELINE.json configs[4] on one GPU of the box: a random-init Janus-Pro-7B-shaped language model (HF LlamaModel,
30 layers, hidden 4096, 32 heads, intermediate 11008, vocab 102400 -- the public Janus-Pro-7B config) followed by
(a) the reference formulation of the head (PyTorch vision_head on every position + get_batch_logps + simpo_loss,
ospo/wrapper/train.py:345-445) and (b) the fused head behind patch_train_wrapper.  Reports the SimPO training-step
time of both, the head's share of the step and the parity of loss / log-probs on the same backbone output.
One process = one GPU's share of config 5 (16 pairs per GPU, configs/step5.yaml:23); the 8-GPU figure is this times 8
plus DDP's all-reduce of the backbone gradients, which is outside the head's path."""
import json
import os
import sys
import time
import types
from pathlib import Path

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import head_oracle as O  # noqa: E402  (checker: the reference formulation of the loss)
from ospo_b200 import patch_train_wrapper  # noqa: E402
from transformers import LlamaConfig, LlamaModel  # noqa: E402

dev = torch.device("cuda:0")
PAIRS = int(os.environ.get("PAIRS", "16"))
LAYERS = int(os.environ.get("LAYERS", "30"))
L, T, H, V = 24, 576, 4096, 16384
torch.manual_seed(0)
cfg = LlamaConfig(hidden_size=H, intermediate_size=11008, num_hidden_layers=LAYERS, num_attention_heads=32,
                  num_key_value_heads=32, vocab_size=102400, max_position_embeddings=16384)
cfg.output_hidden_states = True  # train.py:50
with torch.device(dev):
    backbone = LlamaModel(cfg).to(torch.bfloat16)
backbone.gradient_checkpointing_enable()
backbone.train()
head = O.make_head(H, H, V, seed=5).to(torch.bfloat16).to(dev)
for p in head.parameters():
    p.requires_grad_(False)       # configs/step5.yaml:59-66: gen_head frozen, language model trainable

model = torch.nn.Module()
model.language_model = torch.nn.Module()
model.language_model.model = backbone
model.gen_head = head
g = torch.Generator().manual_seed(1)
emb_c = (torch.randn(PAIRS, L + T, H, generator=g) * 0.02).to(torch.bfloat16).to(dev)
emb_r = (torch.randn(PAIRS, L + T, H, generator=g) * 0.02).to(torch.bfloat16).to(dev)
pad = torch.full((PAIRS, L), -100, dtype=torch.long)
lab_c = torch.cat([pad, torch.randint(0, V, (PAIRS, T), generator=g)], 1).to(dev)
lab_r = torch.cat([pad, torch.randint(0, V, (PAIRS, T), generator=g)], 1).to(dev)
batch = {"chosen_inputs_embeds": emb_c, "chosen_labels": lab_c, "rejected_inputs_embeds": emb_r, "rejected_labels": lab_r}
hp = dict(beta=10.0, gamma_beta_ratio=0.5, label_smoothing=0.0, loss_type="sigmoid", sft_weight=0.0)


def backbone_hidden():
    x = torch.cat([emb_c, emb_r], 0)
    return backbone(inputs_embeds=x, use_cache=False).hidden_states[-1]


def reference_step():
    hidden = backbone_hidden()
    labels = torch.cat([lab_c, lab_r], 0)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = head(hidden)
    logps = O.get_batch_logps(logits, labels, average_log_prob=True)
    losses, _, _ = O.simpo_loss(logps[:PAIRS], logps[PAIRS:], hp["beta"], hp["gamma_beta_ratio"], hp["label_smoothing"],
                                hp["loss_type"])
    loss = losses.mean()
    loss.backward()
    return loss.detach(), logps.detach()


class Wrapper:
    pass


w = Wrapper()
w.model = model
for k, v in hp.items():
    setattr(w, k, v)
w.label_pad_token_id = -100
w.logged = {}
w.log = lambda name, val, **kw: w.logged.__setitem__(name, val)
w.log_dict = lambda d, **kw: w.logged.update(d)
w.concatenated_inputs = types.MethodType(
    lambda self, batch: {"concatenated_inputs_embeds": torch.cat([batch["chosen_inputs_embeds"], batch["rejected_inputs_embeds"]], 0),
                     "concatenated_labels": torch.cat([batch["chosen_labels"], batch["rejected_labels"]], 0)}, w)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        backbone.zero_grad(set_to_none=True)
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


res = {"pairs_per_gpu": PAIRS, "layers": LAYERS, "tokens_per_seq": L + T, "head_frozen": True}
ms_ref, (loss_ref, logps_ref) = timed(reference_step, 2)
gref = backbone.layers[0].self_attn.q_proj.weight.grad.detach().float().clone()
res["reference_head_step_ms"] = ms_ref
res["peak_mem_gb_reference"] = torch.cuda.max_memory_allocated() / 2**30
torch.cuda.reset_peak_memory_stats()

patch_train_wrapper(w, image_span=(L - 1, L - 1 + T))


def fused_step():
    loss = w.get_batch_loss_metrics(batch, "train")
    loss.backward()
    return loss.detach(), None


from ospo_b200 import FusedGenHead, _abi  # noqa: E402

assert isinstance(model.gen_head, FusedGenHead)
launches0 = _abi.load().ospo_head_launch_count()
ms_fused, (loss_fused, _) = timed(fused_step, 2)
res["fused_head_kernel_launches"] = int(_abi.load().ospo_head_launch_count() - launches0)
gfused = backbone.layers[0].self_attn.q_proj.weight.grad.detach().float().clone()
res["fused_head_step_ms"] = ms_fused
res["peak_mem_gb_fused"] = torch.cuda.max_memory_allocated() / 2**30


# head-only times on the same hidden states (what the head contributes to the step)
def backbone_only():
    hidden = backbone_hidden()
    hidden.float().mean().backward()
    return None


ms_bb, _ = timed(backbone_only, 2)
res["backbone_only_step_ms"] = ms_bb
res["head_share_reference"] = max(0.0, (ms_ref - ms_bb) / ms_ref)
res["head_share_fused"] = max(0.0, (ms_fused - ms_bb) / ms_fused)
res["pairs_per_s_reference"] = PAIRS / (ms_ref / 1e3)
res["pairs_per_s_fused"] = PAIRS / (ms_fused / 1e3)
res["loss_reference"], res["loss_fused"] = float(loss_ref), float(loss_fused)
res["loss_rel_err"] = abs(float(loss_ref) - float(loss_fused)) / max(1e-9, abs(float(loss_ref)))
res["chosen_logps_fused"] = float(w.logged["train/logps/chosen"]) if "train/logps/chosen" in w.logged else None
res["chosen_logps_reference"] = float(logps_ref[:PAIRS].mean())
res["rejected_logps_fused"] = float(w.logged["train/logps/rejected"]) if "train/logps/rejected" in w.logged else None
res["rejected_logps_reference"] = float(logps_ref[PAIRS:].mean())
res["backbone_grad_rel_err"] = float((gfused - gref).norm() / gref.norm().clamp_min(1e-20))
print("CONFIG5 " + json.dumps(res), flush=True)
