"""Next row N3: gradient-norm clipping + AdamW for the fused head, on its flat buffers.

The reference trains with ``torch.optim.AdamW`` (ospo/wrapper/train.py:108-115; lr 4e-5, betas (0.9, 0.95),
weight_decay 0, eps 1e-8 in configs/step5.yaml:37-43) under Lightning's ``gradient_clip_val = 1.0``
(``torch.nn.utils.clip_grad_norm_``, ospo/utils/train.py:30,50).  When the head is trainable
(train.py:206-208) its four parameters are 83.9 M elements whose gradient the fused backward already leaves
in one flat fp32 buffer (dW2 | dW1 | db2 | db1, all-reduced once).  ``FusedHeadAdamW`` keeps the master
parameters and both moments in the same layout, so a step is two streaming kernels (``ospo_head_grad_sqnorm``,
``ospo_head_adamw_step``) that also refresh the bf16 GEMM operands -- no per-parameter foreach kernels, no
re-staging cast before the next forward.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .head import FusedGenHead, _HeadParams


class FusedHeadAdamW:
    """``torch.optim.AdamW`` semantics for ``FusedGenHead``'s parameters (same update, same hyper-parameters).

    ``step(other_sqnorm=...)`` clips like ``clip_grad_norm_(all_params, max_norm)``: pass the squared gradient
    norm of every *other* parameter being clipped together with the head (a device scalar) and read the total
    back from ``last_total_norm``; with ``max_norm <= 0`` nothing is clipped.
    """

    def __init__(self, head: FusedGenHead, lr: float = 4e-5, betas=(0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_norm: float = 1.0):
        self.head = head
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, tuple(betas), eps, weight_decay, max_norm
        self.step_count = 0
        self.last_total_norm: Optional[torch.Tensor] = None
        H, E, V = head.n_embed, head.image_token_embed, head.image_token_size
        self._dims = (H, E, V)
        W2, W1 = head.vision_head.weight, head.output_mlp_projector.weight
        B2, B1 = head.vision_head.bias, head.output_mlp_projector.bias
        if not W2.is_cuda:
            raise RuntimeError("FusedHeadAdamW needs the head on a B200 (no CPU path)")
        n = ops.flat_grad_numel(H, E, V)
        dev = W2.device
        with torch.no_grad():
            self.params = torch.empty(n, dtype=torch.float32, device=dev)       # fp32 master, layout W2 | W1 | b2 | b1
            views = ops.split_flat_grads(self.params, H, E, V)                  # (W2, W1, b2, b1) views
            for v, p in zip(views, (W2, W1, B2, B1)):
                v.copy_(p.detach().to(torch.float32))
            self.shadow = torch.empty(V * E + E * H, dtype=torch.bfloat16, device=dev)   # bf16 W2 | W1 for the GEMMs
            self.shadow[:V * E].view(V, E).copy_(views[0])
            self.shadow[V * E:].view(E, H).copy_(views[1])
            # the module's parameters become views of the flat buffers: an update is visible without a copy
            for v, p, sh in zip(views, (W2, W1, B2, B1), (self.shadow[:V * E].view(V, E), self.shadow[V * E:].view(E, H),
                                                        None, None)):
                if p.dtype == torch.float32:
                    p.data = v
                elif p.dtype == torch.bfloat16 and sh is not None:
                    p.data = sh
                # other dtypes / bf16 biases are refreshed by a small copy after every step
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._grads = None
        self._install_operands()

    # the kernels' operand cache of the head points at the shadow / master buffers, keyed like _kernel_params() keys it
    def _install_operands(self) -> None:
        H, E, V = self._dims
        head = self.head
        W2v, W1v, b2v, b1v = ops.split_flat_grads(self.params, H, E, V)
        head._cache = _HeadParams(self.shadow[V * E:].view(E, H), b1v, self.shadow[:V * E].view(V, E), b2v)
        ts = (head.output_mlp_projector.weight, head.output_mlp_projector.bias, head.vision_head.weight,
              head.vision_head.bias)
        head._cache_key = tuple((t.data_ptr(), t._version, t.dtype, t.device) for t in ts)
        head._packed = None        # the decode kernel's packed copy is stale after an update

    def _flat_grads(self, use_last_backward: bool) -> torch.Tensor:
        H, E, V = self._dims
        head = self.head
        # The fused backward's own buffer holds ONE micro-batch's gradient (already all-reduced).  It is the gradient of
        # the step only when exactly one fused backward ran since the last step; with gradient accumulation
        # (accumulate_grad_batches > 1, ospo/utils/train.py:31,51) autograd has summed the micro-batches into .grad, and
        # that sum is gathered instead.
        if use_last_backward and head._flat is not None and head._bwd_count == 1:
            return head._flat
        if self._grads is None:
            self._grads = torch.empty_like(self.params)
        views = ops.split_flat_grads(self._grads, H, E, V)
        for v, p in zip(views, (head.vision_head.weight, head.output_mlp_projector.weight, head.vision_head.bias,
                                head.output_mlp_projector.bias)):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        return self._grads

    @torch.no_grad()
    def step(self, other_sqnorm: Optional[torch.Tensor] = None, use_last_backward: bool = False,
             lr: Optional[float] = None) -> None:
        """one optimizer step.  ``use_last_backward`` reads the flat buffer the last fused backward wrote when exactly
        one fused backward ran since the previous step (no gather); after several micro-batches (gradient
        accumulation) or without the flag the parameters' accumulated ``.grad`` are gathered.  ``lr`` overrides the
        learning rate for this step (schedulers)."""
        g = self._flat_grads(use_last_backward)
        self.head._bwd_count = 0
        total = None
        if self.max_norm > 0:
            total = ops.grad_sqnorm_impl(g)
            if other_sqnorm is not None:
                total = total + other_sqnorm.to(torch.float32).reshape(1)
            self.last_total_norm = total.sqrt()
        self.step_count += 1
        ops.adamw_step_impl(g, self.params, self.exp_avg, self.exp_avg_sq, self.step_count,
                            self.lr if lr is None else lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                            self.max_norm, total, self.shadow)
        H, E, V = self._dims
        head = self.head
        views = ops.split_flat_grads(self.params, H, E, V)
        for v, p in zip(views, (head.vision_head.weight, head.output_mlp_projector.weight, head.vision_head.bias,
                                head.output_mlp_projector.bias)):
            if p.data_ptr() != v.data_ptr() and not (p.dtype == torch.bfloat16 and p.dim() == 2):
                p.data.copy_(v)        # parameters that are not views of the flat buffers (e.g. bf16 biases)
        self._install_operands()

    # ---- checkpoint / resume (Lightning saves optimizer_states next to the model's state_dict) ----------------
    def state_dict(self) -> dict:
        """step count, both moments and the fp32 MASTER parameters (flat), and the hyper-parameters.  The master copy
        is part of the state: with bf16 module parameters the module's own ``state_dict`` only holds their bf16
        rounding, and a resume from that alone would not continue the same trajectory."""
        return {"step": self.step_count, "exp_avg": self.exp_avg.detach().clone(),
                "exp_avg_sq": self.exp_avg_sq.detach().clone(), "params": self.params.detach().clone(),
                "layout": "W2|W1|b2|b1", "dims": self._dims,
                "hyper": {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                          "max_norm": self.max_norm}}

    @torch.no_grad()
    def load_state_dict(self, state: dict) -> None:
        if tuple(state["dims"]) != tuple(self._dims) or state["exp_avg"].numel() != self.exp_avg.numel():
            raise ValueError("optimizer state was saved for a head of another shape")
        self.step_count = int(state["step"])
        self.exp_avg.copy_(state["exp_avg"].to(self.exp_avg.device, torch.float32))
        self.exp_avg_sq.copy_(state["exp_avg_sq"].to(self.exp_avg.device, torch.float32))
        h = state.get("hyper", {})
        self.lr, self.betas = h.get("lr", self.lr), tuple(h.get("betas", self.betas))
        self.eps, self.weight_decay = h.get("eps", self.eps), h.get("weight_decay", self.weight_decay)
        self.max_norm = h.get("max_norm", self.max_norm)
        if "params" in state:
            # exact resume: restore the fp32 masters, then push them into the module (views need nothing; parameters
            # that are not views -- e.g. bf16 biases -- get a copy) and refresh the bf16 GEMM operands
            self.params.copy_(state["params"].to(self.params.device, torch.float32))
            H, E, V = self._dims
            head = self.head
            views = ops.split_flat_grads(self.params, H, E, V)
            self.shadow[:V * E].view(V, E).copy_(views[0])
            self.shadow[V * E:].view(E, H).copy_(views[1])
            for v, p in zip(views, (head.vision_head.weight, head.output_mlp_projector.weight, head.vision_head.bias,
                                    head.output_mlp_projector.bias)):
                if p.data_ptr() != v.data_ptr() and not (p.dtype == torch.bfloat16 and p.dim() == 2):
                    p.data.copy_(v)
            self._install_operands()
        else:
            self.resync_from_module()   # older checkpoints: rebuild the masters from the module's parameters

    @torch.no_grad()
    def resync_from_module(self) -> None:
        """call after the module's parameters were overwritten from outside (``load_state_dict`` on the model):
        refreshes the fp32 master copy where a parameter is not a view of it, and the bf16 operands"""
        H, E, V = self._dims
        head = self.head
        views = ops.split_flat_grads(self.params, H, E, V)
        for v, p in zip(views, (head.vision_head.weight, head.output_mlp_projector.weight, head.vision_head.bias,
                                head.output_mlp_projector.bias)):
            if p.data_ptr() != v.data_ptr():
                v.copy_(p.detach().to(torch.float32))
        self.shadow[:V * E].view(V, E).copy_(views[0])
        self.shadow[V * E:].view(E, H).copy_(views[1])
        self._install_operands()

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.head._bwd_count = 0
        for p in self.head.parameters():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
