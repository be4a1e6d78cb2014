"""CFG image-token generation loop with the fused head (mirror of
``JanusProImageGenWrapper.generate_image``, ospo/wrapper/image_generation.py:143-171, and its copy
``JanusProTestWrapper.generate_batch``, ospo/inference.py:140-163).  The backbone and
``prepare_gen_img_embeds`` stay the caller's PyTorch modules; only lines :156-164 change."""
from __future__ import annotations

from typing import Callable, Optional

import torch

IMAGE_TOKEN_NUM_PER_IMAGE = 576  # ospo/constant.py


@torch.inference_mode()
def generate_image_tokens(gen_head, backbone_step: Callable, prepare_gen_img_embeds: Callable,
                          inputs_embeds: torch.Tensor, attention_masks: torch.Tensor, *, cfg_weight: float = 5.0,
                          temperature: float = 1.0, image_token_num_per_image: int = IMAGE_TOKEN_NUM_PER_IMAGE,
                          uniforms: Optional[torch.Tensor] = None, greedy: bool = False) -> torch.Tensor:
    """inputs_embeds [2P, L, D] (row 2k conditional, 2k+1 unconditional, :132-141) -> tokens [P, n] int32.

    backbone_step(inputs_embeds, attention_mask, past_key_values) -> (last_hidden_state [2P, l, H], past)
    uniforms: optional [n, P] fp32, one row per decode step (else drawn from torch's CUDA generator).
    """
    P = inputs_embeds.shape[0] // 2
    generated = torch.zeros((P, image_token_num_per_image), dtype=torch.int, device=inputs_embeds.device)  # :147
    past = None
    for i in range(image_token_num_per_image):                                                          # :149
        hidden_states, past = backbone_step(inputs_embeds, attention_masks, past)                       # :150-154
        u = None if uniforms is None else uniforms[i]
        if hasattr(prepare_gen_img_embeds, "from_sampled"):
            # FusedGenImgEmbeds: sampling, id duplication, gen_embed and gen_aligner are one launch chain
            next_token, emb = gen_head.cfg_sample(hidden_states[:, -1, :], cfg_weight, temperature, uniforms=u,
                                                  greedy=greedy, next_embeds=prepare_gen_img_embeds)    # :156-168
            generated[:, i] = next_token                                                                # :164
            inputs_embeds = emb.unsqueeze(dim=1)
        else:
            next_token = gen_head.cfg_sample(hidden_states[:, -1, :], cfg_weight, temperature, uniforms=u,
                                             greedy=greedy)                                             # :156-163
            generated[:, i] = next_token                                                                # :164
            both = torch.stack([next_token, next_token], dim=1).view(-1)                                # :166
            inputs_embeds = prepare_gen_img_embeds(both).unsqueeze(dim=1)                               # :167-168
        new_mask = torch.ones((attention_masks.shape[0], 1), dtype=attention_masks.dtype,
                              device=attention_masks.device)
        attention_masks = torch.cat([attention_masks, new_mask], dim=1)                                 # :170-171
    return generated
