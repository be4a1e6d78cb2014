"""Next row N4: the hand-off of the sampled ids to the VQ decoder and to disk, byte-compatible with the reference
(``JanusProImageGenWrapper.generate_image``, ospo/wrapper/image_generation.py:147,164,174-191).

* token buffer: ``int32 [P, 576]`` (``:147``), one column per decode step (``:164``);
* decoder call: ``gen_vision_model.decode_code(tokens.to(torch.int), shape=[P, 8, 24, 24])`` (``:174``,
  janus/models/vq_model.py:505-508) -- the CNN decoder stays the caller's PyTorch module;
* pixels: ``clip((dec + 1) / 2 * 255, 0, 255)`` in fp32, stored into a uint8 array (C truncation), NHWC (``:175-180``);
* files: one PNG per image at ``save_path_list[i]``, with the reference's fallback name on ``OSError`` (``:184-191``).

``tokens_to_uint8`` does the fp32 arithmetic and the cast on the device, in the reference's operation order, so the
bytes are identical while 4x less data crosses PCIe (uint8 NHWC instead of fp32 NCHW).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np
import torch

IMG_SIZE, PATCH_SIZE, CODE_DIM = 384, 16, 8  # ospo/constant.py:1-4, image_generation.py:116-117


@torch.no_grad()
def tokens_to_uint8(decode_code: Callable, generated_tokens: torch.Tensor, img_size: int = IMG_SIZE,
                    patch_size: int = PATCH_SIZE) -> np.ndarray:
    """generated_tokens [P, (img_size/patch_size)^2] -> uint8 [P, img_size, img_size, 3]  (image_generation.py:174-180)"""
    P = generated_tokens.shape[0]
    side = img_size // patch_size
    dec = decode_code(generated_tokens.to(dtype=torch.int), shape=[P, CODE_DIM, side, side])
    dec = dec.to(torch.float32)
    px = ((dec + 1) / 2 * 255).clamp_(0, 255)          # same three fp32 operations, same order
    img = px.to(torch.uint8).permute(0, 2, 3, 1).contiguous()   # float -> uint8 truncates, like the numpy store
    out = img.cpu().numpy()
    assert out.shape == (P, img_size, img_size, 3)
    return out


def save_images(images: np.ndarray, save_path_list: Sequence[str]) -> List[str]:
    """image_generation.py:184-191: PNG per image; on OSError fall back to ``longprompt_<idx>.png`` in the cwd.
    Returns the paths actually written."""
    from PIL import Image

    written = []
    for inner_idx, image in enumerate(images):
        path = save_path_list[inner_idx]
        try:
            Image.fromarray(image).save(path)
        except OSError:
            idx_in_path = path.split("_")[1]
            path = f"longprompt_{idx_in_path}"
            Image.fromarray(image).save(path)
        written.append(path)
    return written
