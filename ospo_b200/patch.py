"""Plug-in points for the reference code base.

* ``register_gen_head_cls`` adds ``"vision_head_sm100"`` to the class registry the reference uses to
  build ``gen_head`` (``model_name_to_cls``, janus/models/modeling_vlm.py:54-70; selected by
  ``config.gen_head_config.cls``, :133-145, 210-212).
* ``patch_model`` swaps an already-built ``model.gen_head`` (MultiModalityCausalLM, :210-212) for a
  ``FusedGenHead`` that shares its parameters.
* ``patch_train_wrapper`` re-points ``JanusProTrainWrapper.concatenated_forward`` /
  ``get_batch_loss_metrics`` (ospo/wrapper/train.py:345-372, 399-445) at the fused path while keeping
  their signatures and return values, and ``preprocess_batch`` (:219-279) at a batched form of the VQ encode
  (SURVEY §8f N2).
"""
from __future__ import annotations

import types
from typing import Dict, Optional

import torch

from .head import FusedGenHead, FusedGenImgEmbeds

SM100_CLS_NAME = "vision_head_sm100"


def register_gen_head_cls(modeling_vlm_module) -> None:
    """wrap ``modeling_vlm.model_name_to_cls`` so ``"vision_head_sm100"`` resolves to FusedGenHead"""
    orig = modeling_vlm_module.model_name_to_cls

    def model_name_to_cls(cls_name):
        if SM100_CLS_NAME in cls_name:
            return FusedGenHead
        return orig(cls_name)

    modeling_vlm_module.model_name_to_cls = model_name_to_cls


def patch_model(model: torch.nn.Module, fuse_gen_img_embeds: bool = False, embed_table: bool = False) -> torch.nn.Module:
    """replace ``model.gen_head`` in place; parameters (and their requires_grad flags) are shared.
    ``fuse_gen_img_embeds`` also re-points ``model.prepare_gen_img_embeds`` (modeling_vlm.py:263-264) at the fused
    gen_embed -> gen_aligner kernels (generation only); ``embed_table`` additionally memoises that module over the
    16384 codes (``FusedGenImgEmbeds.build_table``: +134 MB for the 7B model, the decode loop then copies two table
    rows per pair instead of streaming the aligner's weights every step)."""
    if not isinstance(model.gen_head, FusedGenHead):
        model.gen_head = FusedGenHead.from_reference(model.gen_head)
    if fuse_gen_img_embeds and hasattr(model, "gen_embed") and hasattr(model, "gen_aligner"):
        model.prepare_gen_img_embeds = FusedGenImgEmbeds(model.gen_embed, model.gen_aligner)
        if embed_table:
            model.prepare_gen_img_embeds.build_table()
    return model


def batched_preprocess_batch(self, batch, token_cache: Optional[dict] = None):
    """``JanusProTrainWrapper.preprocess_batch`` (ospo/wrapper/train.py:219-279) with the per-sample work batched
    (SURVEY §8f N2).  Same inputs, same output dictionary, same values:

    * the reference VQ-encodes chosen and rejected images one at a time -- 2B batch-1 passes of the CNN encoder
      (:246-261); here all 2B images go through ``gen_vision_model.encode`` in ONE ``[2B, 3, H, W]`` call and the
      code indices (``output[2][2]``, vq_model.py:278-282) are split back per image;
    * ``token_cache`` (a dict, e.g. one per dataset): the images of a dataset item never change, so its token ids are
      kept under ``item_id`` after the first epoch and the encoder is skipped for cached items;
    * the text prompts are embedded in one padded lookup instead of B lookups; padded positions are zero embeddings
      with label -100 exactly as at :229-239;
    * ``prepare_gen_img_embeds`` runs once on the ``[2B, T]`` ids instead of twice.
    """
    item_ids, text_tokens, chosen_image_tensors, rejected_image_tensors = batch
    B = len(item_ids)
    dev, dtype = self.device, self.model.dtype
    # ---- text: one padded embedding lookup, zero rows on the padding (train.py:224-239)
    lens = [int(t.shape[1]) for t in text_tokens]
    max_len = max(lens)
    ids = torch.zeros(B, max_len, dtype=torch.long, device=dev)
    keep = torch.zeros(B, max_len, 1, dtype=dtype, device=dev)
    for i, t in enumerate(text_tokens):
        ids[i, :lens[i]] = t.reshape(-1).to(dev)
        keep[i, :lens[i]] = 1
    text_embeds = self.model.language_model.get_input_embeddings()(ids).to(dtype) * keep
    text_labels = torch.full((B, max_len), -100, dtype=torch.long, device=dev)
    # ---- images: one batched VQ encode for everything that is not cached (train.py:241-264)
    vq = self.model.gen_vision_model
    vq_dtype = next(vq.parameters()).dtype
    tokens = [None] * (2 * B)                       # chosen 0..B-1, rejected B..2B-1
    todo = []
    for i in range(B):
        hit = token_cache.get(item_ids[i]) if token_cache is not None else None
        if hit is not None:
            tokens[i], tokens[B + i] = hit[0].to(dev), hit[1].to(dev)
        else:
            todo.append(i)
    if todo:
        imgs = torch.cat([chosen_image_tensors[i] for i in todo] + [rejected_image_tensors[i] for i in todo], 0)
        out = vq.encode(imgs.to(device=dev, dtype=vq_dtype))
        codes = out[2][2].reshape(2 * len(todo), -1)
        for j, i in enumerate(todo):
            tokens[i], tokens[B + i] = codes[j], codes[len(todo) + j]
            if token_cache is not None:
                token_cache[item_ids[i]] = (codes[j].detach().clone(), codes[len(todo) + j].detach().clone())
    all_tokens = torch.stack(tokens, 0)              # [2B, T]
    img_embeds = self.model.prepare_gen_img_embeds(all_tokens).to(dev)
    chosen_tok, rejected_tok = all_tokens[:B], all_tokens[B:]
    pre = {"item_ids": item_ids}
    pre["chosen_inputs_embeds"] = torch.cat([text_embeds, img_embeds[:B]], dim=1)
    pre["chosen_attention_mask"] = torch.ones(pre["chosen_inputs_embeds"].shape[:2], dtype=torch.long)
    pre["chosen_labels"] = torch.cat([text_labels, chosen_tok], dim=1)
    pre["rejected_inputs_embeds"] = torch.cat([text_embeds, img_embeds[B:]], dim=1)
    pre["rejected_attention_mask"] = torch.ones(pre["rejected_inputs_embeds"].shape[:2], dtype=torch.long)
    pre["rejected_labels"] = torch.cat([text_labels, rejected_tok], dim=1)
    return pre


def patch_train_wrapper(wrapper, image_span=None, process_group=None, batch_vq_encode: bool = True,
                        cache_image_tokens: bool = False):
    """``wrapper``: a JanusProTrainWrapper (or anything with .model.gen_head, .model.language_model.model,
    .concatenated_inputs, the SimPO hyper-parameter attributes and .log/.log_dict).
    ``batch_vq_encode``: re-point ``preprocess_batch`` at :func:`batched_preprocess_batch` (needs
    ``.model.gen_vision_model``); ``cache_image_tokens`` additionally keeps every item's token ids after its first
    encode (``wrapper.image_token_cache``)."""
    patch_model(wrapper.model)
    if batch_vq_encode and hasattr(wrapper.model, "gen_vision_model"):
        wrapper.image_token_cache = {} if cache_image_tokens else None
        wrapper.preprocess_batch = types.MethodType(
            lambda self, batch: batched_preprocess_batch(self, batch, self.image_token_cache), wrapper)

    def concatenated_forward(self, batch: Dict):
        # train.py:345-372, with gen_head + get_batch_logps fused.  The [S, L+T, V] logits are never built, so the
        # two logits entries of the reference 5-tuple are None here; the only thing the reference does with them is
        # the mean it logs at :441-442, which the fused get_batch_loss_metrics below reports from the kernel's own
        # row sums (metrics["logits/chosen"], ["logits/rejected"]).
        concatenated_batch = self.concatenated_inputs(batch=batch)
        len_chosen = batch["chosen_labels"].shape[0]
        outputs = self.model.language_model.model(
            inputs_embeds=concatenated_batch["concatenated_inputs_embeds"], use_cache=False, past_key_values=None)
        hidden_states = outputs.hidden_states[-1]
        labels = concatenated_batch["concatenated_labels"]
        all_logps = self.model.gen_head.logps(hidden_states, labels, average_log_prob=True,
                                              ignore_index=self.label_pad_token_id, image_span=image_span,
                                              process_group=process_group)
        return (all_logps[:len_chosen], all_logps[len_chosen:], None, None, labels[:len_chosen])

    def get_batch_loss_metrics(self, batch: Dict, train_eval: str = "train"):
        # train.py:399-445 in one fused forward (+ backward through autograd)
        prefix = "val" if train_eval == "val" else "train"
        concatenated_batch = self.concatenated_inputs(batch=batch)
        outputs = self.model.language_model.model(
            inputs_embeds=concatenated_batch["concatenated_inputs_embeds"], use_cache=False, past_key_values=None)
        hidden_states = outputs.hidden_states[-1]
        out = self.model.gen_head.simpo(
            hidden_states, concatenated_batch["concatenated_labels"], beta=self.beta,
            gamma_beta_ratio=self.gamma_beta_ratio, label_smoothing=self.label_smoothing, loss_type=self.loss_type,
            sft_weight=self.sft_weight, ignore_index=self.label_pad_token_id, image_span=image_span,
            process_group=process_group)
        if self.sft_weight > 0.0:
            self.log(f"{prefix}/sft_loss", out.metrics["sft_loss"], on_step=True, prog_bar=True, logger=True,
                     sync_dist=True)
        # device scalars: no .cpu() here, so the step does not synchronise
        self.log_dict({f"{prefix}/{k}": v for k, v in out.metrics.items() if "/" in k},
                      on_step=True, prog_bar=True, logger=True, sync_dist=True)
        return out.loss

    def compute_total_grad_norm(self):
        # train.py:459-469 without its per-parameter ``.item()`` syncs: same value
        # (sqrt of the sum of squared parameter-gradient norms), returned as a device scalar
        if hasattr(self.trainer.model, "get_global_grad_norm"):
            grad_norm = self.trainer.model.get_global_grad_norm()
            return grad_norm if grad_norm is not None else 0.0
        return total_grad_norm(self.model.parameters())

    wrapper.concatenated_forward = types.MethodType(concatenated_forward, wrapper)
    wrapper.get_batch_loss_metrics = types.MethodType(get_batch_loss_metrics, wrapper)
    wrapper.compute_total_grad_norm = types.MethodType(compute_total_grad_norm, wrapper)
    return wrapper


def total_grad_norm(parameters):
    """sqrt(sum_p ||p.grad||_2^2) over the parameters that have a gradient, as one device scalar (0.0 if none) --
    the quantity ``JanusProTrainWrapper.compute_total_grad_norm`` logs (train.py:464-469), computed with two
    multi-tensor kernels instead of one ``.item()`` round trip per parameter."""
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if not grads:
        return 0.0
    norms = torch._foreach_norm(grads, 2)
    return torch.linalg.vector_norm(torch.stack([n.to(torch.float32) for n in norms]), 2)
