"""TEST INFRASTRUCTURE -- CPU oracle for OSPO's image-token head path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``ospo_b200/`` does.

It is a plain-PyTorch (CPU) restatement of the reference lines cited on every function (paths relative
to the OSPO repository).  **Parity is pinned**: ``tests/golden/simpo_ref.npz`` and
``tests/golden/cfg_ref.npz`` were produced by executing the reference's own source files
(``tests/golden/make_golden.py``), and ``tests/test_oracle.py`` checks this restatement against them.

The arithmetic itself lives in third-party PyTorch (aten ``addmm``/``gelu``/``_log_softmax``/``gather``/
``log_sigmoid``/``_softmax``/``multinomial``; the reference pins torch==2.0.1, requirements.txt:2).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = Path(__file__).resolve().parent
IMAGE_TOKEN_NUM_PER_IMAGE = 576  # ospo/constant.py:3-4 (384 / 16) ** 2


# --------------------------------------------------------------------------------------------------
# head: janus/models/modeling_vlm.py:36-51
# --------------------------------------------------------------------------------------------------
class VisionHead(torch.nn.Module):
    """``vision_head``: Linear(n_embed -> image_token_embed) -> exact GELU -> Linear(-> image_token_size)."""

    def __init__(self, n_embed: int, image_token_embed: int, image_token_size: int):
        super().__init__()
        self.output_mlp_projector = torch.nn.Linear(n_embed, image_token_embed)  # :39-41
        self.vision_activation = torch.nn.GELU()                                 # :42
        self.vision_head = torch.nn.Linear(image_token_embed, image_token_size)  # :43-45

    def forward(self, x):  # :47-51
        x = self.output_mlp_projector(x)
        x = self.vision_activation(x)
        x = self.vision_head(x)
        return x


def make_head(H: int, E: int, V: int, seed: int, dtype=torch.float32, w2_gain: float = 1.0) -> VisionHead:
    """default nn.Linear init under a fixed seed ("identical random-init weights")"""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    head = VisionHead(H, E, V)
    torch.random.set_rng_state(g)
    if w2_gain != 1.0:
        with torch.no_grad():
            head.vision_head.weight.mul_(w2_gain)
    return head.to(dtype)


# --------------------------------------------------------------------------------------------------
# SimPO: ospo/wrapper/train.py
# --------------------------------------------------------------------------------------------------
def get_batch_logps(logits: torch.Tensor, labels: torch.Tensor, average_log_prob: bool = True,
                    label_pad_token_id: int = -100, return_per_token: bool = False):
    """train.py:375-396"""
    if logits.shape[:-1] != labels.shape:
        raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
    labels = labels[:, 1:].clone()                     # :385
    logits = logits[:, :-1, :]                         # :386
    loss_mask = labels != label_pad_token_id           # :387
    labels[labels == label_pad_token_id] = 0           # :389
    # :391 -- under the reference's autocast, log_softmax runs in float32
    per_token_logps = torch.gather(logits.float().log_softmax(-1), dim=2, index=labels.unsqueeze(2)).squeeze(2)
    if average_log_prob:
        out = (per_token_logps * loss_mask).sum(-1) / loss_mask.sum(-1)   # :394
    else:
        out = (per_token_logps * loss_mask).sum(-1)                       # :396
    return (out, per_token_logps, loss_mask) if return_per_token else out


def simpo_loss(chosen_logps: torch.Tensor, rejected_logps: torch.Tensor, beta: float, gamma_beta_ratio: float,
               label_smoothing: float = 0.0, loss_type: str = "sigmoid"):
    """train.py:317-342"""
    pi_logratios = chosen_logps - rejected_logps
    logits = pi_logratios - gamma_beta_ratio
    if loss_type == "sigmoid":
        losses = (-F.logsigmoid(beta * logits) * (1 - label_smoothing)
                  - F.logsigmoid(-beta * logits) * label_smoothing)
    elif loss_type == "hinge":
        losses = torch.relu(1 - beta * logits)
    else:
        raise ValueError(f"Unknown loss type: {loss_type}. Should be one of ['sigmoid', 'hinge']")
    chosen_rewards = beta * chosen_logps.detach()
    rejected_rewards = beta * rejected_logps.detach()
    return losses, chosen_rewards, rejected_rewards


def simpo_step(head: VisionHead, hidden_chosen: torch.Tensor, hidden_rejected: torch.Tensor,
               labels_chosen: torch.Tensor, labels_rejected: torch.Tensor, *, beta: float = 10.0,
               gamma_beta_ratio: float = 0.5, label_smoothing: float = 0.0, sft_weight: float = 0.0,
               loss_type: str = "sigmoid", backward: bool = False) -> Dict[str, torch.Tensor]:
    """concatenated_forward (train.py:345-372) on given last-hidden-states + get_batch_loss_metrics (:399-445).

    hidden_*: [B, L+T, H]; labels_*: [B, L+T] with -100 on the text positions.  The dtype of ``head`` /
    ``hidden_*`` selects the mode: float32 (config 1) or bfloat16 (Linear/GELU in bf16, log-softmax in fp32,
    i.e. the reference's bf16 GPU semantics).
    """
    B = labels_chosen.shape[0]                                      # len_chosen :349
    hidden = torch.cat([hidden_chosen, hidden_rejected], dim=0)     # concatenated_inputs :282-314
    labels = torch.cat([labels_chosen, labels_rejected], dim=0)
    if backward:
        hidden = hidden.detach().clone().requires_grad_(True)
        head.zero_grad()
    all_logits = head(hidden)                                       # :357
    all_logps, per_tok, mask = get_batch_logps(all_logits, labels, average_log_prob=True, return_per_token=True)
    chosen_logps, rejected_logps = all_logps[:B], all_logps[B:]     # :364-365
    chosen_logits, rejected_logits = all_logits[:B], all_logits[B:]
    losses, chosen_rewards, rejected_rewards = simpo_loss(chosen_logps, rejected_logps, beta, gamma_beta_ratio,
                                                          label_smoothing, loss_type)   # :414-417
    loss = losses.mean()                                            # :419
    sft_loss = torch.zeros(())
    logged_chosen_logits = chosen_logits
    if sft_weight > 0.0:                                            # :421-430
        pl = chosen_logits[..., :-1, :].contiguous()
        cl = labels_chosen[..., 1:].clone()
        sft_loss = torch.nn.CrossEntropyLoss()(pl.view(-1, pl.shape[-1]).float(), cl.view(-1))
        loss = sft_weight * sft_loss + loss
        # reference quirk: :422 rebinds policy_chosen_logits to the [:-1] slice, so the metric logged at
        # :442 is the mean of the sliced tensor whenever the SFT term is on
        logged_chosen_logits = pl
    out = dict(
        loss=loss, losses=losses, chosen_logps=chosen_logps, rejected_logps=rejected_logps,
        chosen_rewards=chosen_rewards, rejected_rewards=rejected_rewards, sft_loss=sft_loss,
        per_token_logps=per_tok, loss_mask=mask,
        reward_accuracy=(chosen_rewards > rejected_rewards).float().mean(),      # :432
        reward_margin=(chosen_rewards - rejected_rewards).mean(),                # :433
        logits_chosen_mean=logged_chosen_logits.detach().float().mean(),         # :442
        logits_rejected_mean=rejected_logits.detach().float().mean(),            # :441
        # the fused head only evaluates the unmasked rows; their mean logit is reported beside the above
        logits_chosen_valid_mean=_valid_row_mean(chosen_logits.detach(), mask[:B]),
        logits_rejected_valid_mean=_valid_row_mean(rejected_logits.detach(), mask[B:]),
    )
    if backward:
        loss.backward()
        out.update(
            dx=hidden.grad,
            dW1=head.output_mlp_projector.weight.grad, db1=head.output_mlp_projector.bias.grad,
            dW2=head.vision_head.weight.grad, db2=head.vision_head.bias.grad,
        )
    return out


def _valid_row_mean(logits: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    rows = logits[:, :-1, :][mask]          # rows whose (shifted) label is not -100
    return rows.float().mean() if rows.numel() else torch.zeros(())


def analytic_row_coefficients(chosen_logps, rejected_logps, n_rows_per_seq: int, *, beta, gamma_beta_ratio,
                              label_smoothing=0.0, loss_type="sigmoid"):
    """g_b = d loss / d chosen_logps[b]; per-row coefficient g/n (used to cross-check autograd)."""
    B = chosen_logps.shape[0]
    z = (chosen_logps - rejected_logps) - gamma_beta_ratio
    if loss_type == "sigmoid":
        g = -(beta / B) * ((1 - label_smoothing) * torch.sigmoid(-beta * z) - label_smoothing * torch.sigmoid(beta * z))
    else:
        g = -(beta / B) * ((1 - beta * z) > 0).float()
    return g / n_rows_per_seq, -g / n_rows_per_seq


# --------------------------------------------------------------------------------------------------
# CFG decode tail: ospo/wrapper/image_generation.py:156-164 (== ospo/inference.py:147-155)
# --------------------------------------------------------------------------------------------------
def cfg_probs(logits: torch.Tensor, cfg_weight: float, temperature: float) -> torch.Tensor:
    """image_generation.py:157-161 with the reference's tensor dtypes: elementwise ops in the dtype of
    ``logits`` (bf16 on the reference path => rounding after each op), softmax in float32 (CUDA autocast)."""
    logit_cond = logits[0::2, :]                                     # :157
    logit_uncond = logits[1::2, :]                                   # :158
    merged = logit_uncond + cfg_weight * (logit_cond - logit_uncond)  # :160
    return torch.softmax((merged / temperature).float(), dim=-1)     # :161


def cfg_merged(logits: torch.Tensor, cfg_weight: float, temperature: float) -> torch.Tensor:
    logit_cond = logits[0::2, :]
    logit_uncond = logits[1::2, :]
    return (logit_uncond + cfg_weight * (logit_cond - logit_uncond)) / temperature


def decode_step_reference(head: VisionHead, hidden_last: torch.Tensor, cfg_weight: float, temperature: float,
                          generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """one iteration of the generate loop, image_generation.py:156-163, including torch.multinomial"""
    logits = head(hidden_last)                                       # :156
    probs = cfg_probs(logits, cfg_weight, temperature)
    next_token = torch.multinomial(probs, num_samples=1, generator=generator)   # :163
    return next_token.squeeze(-1), probs


# --------------------------------------------------------------------------------------------------
# next row N1: prepare_gen_img_embeds (janus/models/modeling_vlm.py:263-264)
# --------------------------------------------------------------------------------------------------
class GenAligner(torch.nn.Module):
    """``MlpProjector`` with ``projector_type='mlp_gelu', depth=2`` (janus/models/projector.py:39-45, 86):
    layers = Sequential(Linear(input_dim, n_embed), GELU(), Linear(n_embed, n_embed))"""

    def __init__(self, input_dim: int, n_embed: int):
        super().__init__()
        self.layers = torch.nn.Sequential(torch.nn.Linear(input_dim, n_embed), torch.nn.GELU(),
                                          torch.nn.Linear(n_embed, n_embed))

    def forward(self, x):
        return self.layers(x)


def prepare_gen_img_embeds(gen_embed: torch.nn.Embedding, gen_aligner: torch.nn.Module, image_ids: torch.Tensor):
    """modeling_vlm.py:263-264"""
    return gen_aligner(gen_embed(image_ids))


# ---- deterministic inverse-CDF sampler (oracle/cfg_sample.c) ---------------------------------------
_clib = None


def build_c_oracle(force: bool = False) -> Path:
    """gcc-compile oracle/cfg_sample.c into oracle/_build/liboracle.so (test infrastructure)."""
    out_dir = _HERE / "_build"
    out = out_dir / "liboracle.so"
    src = _HERE / "cfg_sample.c"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        out_dir.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(out),
                        str(src), "-lm"], check=True)
    return out


def _c():
    global _clib
    if _clib is None:
        lib = C.CDLL(str(build_c_oracle()))
        lib.ospo_oracle_exp_det.argtypes = [C.c_float]
        lib.ospo_oracle_exp_det.restype = C.c_float
        lib.ospo_oracle_cfg_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                               C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ospo_oracle_cfg_sample.restype = C.c_int
        _clib = lib
    return _clib


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.bfloat16).contiguous().cpu().view(torch.int16).numpy().view(np.uint16)


def bits_to_bf16(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


def cfg_sample_det(logits_bf16: torch.Tensor, cfg_weight: float, temperature: float,
                   uniforms: Optional[torch.Tensor], merge_mode: int = 0, greedy: bool = False):
    """inverse-CDF (or greedy) sampling on bf16 logits [2P, V]; returns (ids[P], merged[P,V], weights[P,V], Z[P])."""
    bits = np.ascontiguousarray(bf16_bits(logits_bf16))
    twoP, V = bits.shape
    P = twoP // 2
    u = np.zeros(P, np.float32) if uniforms is None else np.ascontiguousarray(uniforms.detach().cpu().numpy(), np.float32)
    ids = np.zeros(P, np.int64)
    merged = np.zeros((P, V), np.float32)
    weights = np.zeros((P, V), np.float32)
    Z = np.zeros(P, np.float32)
    rc = _c().ospo_oracle_cfg_sample(bits.ctypes.data, P, V, float(cfg_weight), float(temperature), int(merge_mode),
                                     u.ctypes.data, int(greedy), ids.ctypes.data, merged.ctypes.data,
                                     weights.ctypes.data, Z.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle sampler failed: {rc}")
    return torch.from_numpy(ids), torch.from_numpy(merged), torch.from_numpy(weights), torch.from_numpy(Z)


def exp_det(x: float) -> float:
    return float(_c().ospo_oracle_exp_det(float(x)))


# --------------------------------------------------------------------------------------------------
# synthetic inputs: everything from a CPU generator so oracle and kernel see identical bits
# --------------------------------------------------------------------------------------------------
def synthetic_simpo_batch(B: int, T: int, L: int, H: int, V: int, seed: int, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    hc = torch.randn(B, L + T, H, generator=g).to(dtype)
    hr = torch.randn(B, L + T, H, generator=g).to(dtype)
    ic = torch.randint(0, V, (B, T), generator=g)
    ir = torch.randint(0, V, (B, T), generator=g)
    pad = torch.full((B, L), -100, dtype=torch.long)
    return hc, hr, torch.cat([pad, ic], 1), torch.cat([pad, ir], 1)


# --------------------------------------------------------------------------------------------------
# next row N3: clip_grad_norm_ + AdamW (restated; the reference uses torch's own implementations:
# ospo/utils/train.py:30,50 gradient_clip_val = 1.0 and ospo/wrapper/train.py:108-115 torch.optim.AdamW)
# --------------------------------------------------------------------------------------------------
def clip_adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr=4e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0,
                    max_norm=1.0, other_sqnorm=0.0):
    """One ``clip_grad_norm_(max_norm)`` + ``torch.optim.AdamW`` step on fp32 tensors, written out operation by
    operation in the order PyTorch 2.x performs them (torch/nn/utils/clip_grad.py; torch/optim/adamw.py
    ``_single_tensor_adamw``).  Lists of tensors; updated in place; returns the total gradient norm."""
    sq = sum(float((g.double() ** 2).sum()) for g in grads) + float(other_sqnorm)
    total_norm = sq ** 0.5
    if max_norm > 0:
        coef = min(1.0, max_norm / (total_norm + 1e-6))
        grads = [g * torch.tensor(coef, dtype=torch.float32) for g in grads]
    b1, b2 = betas
    bias1 = 1 - b1 ** step
    bias2_sqrt = (1 - b2 ** step) ** 0.5
    step_size = lr / bias1
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        p.mul_(1 - lr * weight_decay)
        m.add_((g - m) * (1 - b1))                       # lerp_(g, 1 - beta1)
        v.mul_(b2).add_(g * g * (1 - b2))                # addcmul_(g, g, value = 1 - beta2)
        denom = (v.sqrt() / bias2_sqrt).add_(eps)
        p.add_(m / denom * (-step_size))                 # addcdiv_(m, denom, value = -step_size)
    return total_norm
