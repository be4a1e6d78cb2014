"""B200-native (sm_100a) image-token head for OSPO / Janus-Pro: the gen_head MLP with fused SimPO
forward+backward and CFG merge+sample, behind the C ABI in include/ospo_head.h."""
from . import _abi  # noqa: F401
from .head import FusedGenHead, FusedGenImgEmbeds, SimpoOutput, cfg_merge_sample  # noqa: F401
from .optim import FusedHeadAdamW  # noqa: F401
from .patch import patch_model, patch_train_wrapper, register_gen_head_cls  # noqa: F401

__all__ = ["FusedGenHead", "FusedGenImgEmbeds", "FusedHeadAdamW", "SimpoOutput", "cfg_merge_sample", "patch_model", "patch_train_wrapper",
           "register_gen_head_cls"]
