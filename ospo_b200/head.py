"""``FusedGenHead`` -- drop-in replacement for Janus-Pro's ``vision_head`` (the ``gen_head`` module).

Reference: ``janus/models/modeling_vlm.py:36-51`` (module), ``:54-70, 210-212`` (class registry /
construction), consumers ``ospo/wrapper/train.py:345-445`` (SimPO) and
``ospo/wrapper/image_generation.py:156-164`` (CFG merge + sample).

The parameters keep the reference names and ``[out, in]`` layouts
(``output_mlp_projector.{weight,bias}``, ``vision_head.{weight,bias}``) so HF / Lightning checkpoints
load with ``strict=True``.  All compute runs in the sm_100a library behind ``include/ospo_head.h``;
on anything else the calls raise.
"""
from __future__ import annotations

import os
from typing import NamedTuple, Optional, Tuple

import torch

from . import _abi, ops
from . import dist as _dist
from .dist import allreduce_mean_

IGNORE_INDEX = -100  # tokenizer.label_pad_token_id, configs/step5.yaml:73


class SimpoOutput(NamedTuple):
    loss: torch.Tensor               # scalar, differentiable          train.py:419 (+ :428)
    chosen_logps: torch.Tensor       # [B]                             train.py:364
    rejected_logps: torch.Tensor     # [B]                             train.py:365
    losses: torch.Tensor             # [B]                             train.py:329-334
    chosen_rewards: torch.Tensor     # [B]                             train.py:339
    rejected_rewards: torch.Tensor   # [B]                             train.py:340
    per_token_logps: torch.Tensor    # [rows]  log-prob of each unmasked image token
    metrics: dict                    # 0-d device tensors, keys as logged at train.py:432-443


def _rows_from_labels(hidden: torch.Tensor, labels: torch.Tensor, ignore_index: int,
                      image_span: Optional[Tuple[int, int]]):
    """Apply the shift of get_batch_logps (train.py:385-387) and keep only the unmasked rows.

    hidden [S, L, H], labels [S, L]  ->  x (either the gathered rows [N, H], a differentiable copy, or -- zero-copy
    -- ``hidden`` itself plus a (rows per sequence, first row) segment description), targets [N] int64,
    seq_offsets [S+1] int64 (device), seg.
    ``image_span=(start, stop)`` promises that exactly positions start..stop-1 of the *shifted* sequence are
    unmasked in every sequence (the OSPO layout: L text positions then 576 image tokens); it avoids the
    device->host sync of a data-dependent gather.
    """
    S, L, H = hidden.shape
    lab = labels[:, 1:]
    hid = hidden[:, :-1, :]
    if image_span is not None:
        a, b = image_span
        n = b - a
        targets = lab[:, a:b].reshape(-1)
        seq_off = torch.arange(0, (S + 1) * n, n, dtype=torch.int64, device=hidden.device)
        if n % 64 == 0 and hidden.dtype == torch.bfloat16 and hidden.is_contiguous() and hidden.is_cuda:
            # zero-copy: the kernels read (and write dX into) the [S, L, H] tensor through a row-segmented
            # TMA view -- rows a..b-1 of every sequence (the shift only drops the last position)
            return hidden, targets.contiguous(), seq_off, (n, a)
        x_rows = hid[:, a:b, :].reshape(-1, H)
        return x_rows, targets.contiguous(), seq_off, (0, 0)
    mask = lab != ignore_index
    counts = mask.sum(dim=1)
    seq_off = torch.zeros(S + 1, dtype=torch.int64, device=hidden.device)
    seq_off[1:] = torch.cumsum(counts, 0)
    x_rows = hid[mask]          # [N, H], differentiable gather (one host sync for N)
    targets = lab[mask]
    return x_rows, targets.contiguous(), seq_off, (0, 0)


class _HeadParams(NamedTuple):
    w1: torch.Tensor  # bf16 [E, H]
    b1: torch.Tensor  # fp32 [E]
    w2: torch.Tensor  # bf16 [V, E]
    b2: torch.Tensor  # fp32 [V]


class _SimpoFn(torch.autograd.Function):
    """loss = SimPO(head(x_rows)); everything else is returned detached."""

    @staticmethod
    def forward(ctx, x_rows, W1, B1, W2, B2, head, targets, seq_off, hp, group, seg):
        p = head._kernel_params()
        xb = x_rows.detach().to(torch.bfloat16).contiguous()
        need_bwd = any(ctx.needs_input_grad[:5])   # grad mode is off inside forward(); ask autograd instead
        beta, gbr, ls, sftw, lt = hp
        (scalars, seq_logps, losses, crew, rrew, row_logps, row_lse, row_ref, grad_seq, pre, act, gspill) = \
            ops.simpo_fwd_impl(xb, p.w1, p.b1, p.w2, p.b2, targets, seq_off, beta, gbr, ls, sftw, lt, need_bwd, seg[0],
                               seg[1])
        ctx.head, ctx.hp, ctx.group, ctx.seg = head, hp, group, seg
        ctx.x_dtype = x_rows.dtype
        ctx.need_dx = ctx.needs_input_grad[0]
        ctx.need_dw = any(ctx.needs_input_grad[1:5])
        ctx.param_dtypes = (W1.dtype, B1.dtype, W2.dtype, B2.dtype)
        if need_bwd:
            ctx.save_for_backward(xb, p.w1, p.b1, p.w2, p.b2, targets, seq_off, scalars, pre, act, gspill, row_lse,
                                  row_ref, grad_seq)
        loss = scalars[_abi.SC_LOSS].clone()
        ctx.mark_non_differentiable(scalars, seq_logps, losses, crew, rrew, row_logps)
        ctx.set_materialize_grads(False)  # no zero tensors for the six outputs that carry no gradient
        return loss, scalars, seq_logps, losses, crew, rrew, row_logps

    @staticmethod
    def backward(ctx, grad_loss, *_):
        if grad_loss is None:
            return (None,) * 11
        xb, w1, b1, w2, b2, targets, seq_off, scalars, pre, act, gspill, row_lse, row_ref, grad_seq = ctx.saved_tensors
        head = ctx.head
        H, E, V = head.n_embed, head.image_token_embed, head.image_token_size
        flat = head._flat_grad_buffer(ctx.group) if ctx.need_dw else torch.empty(0, dtype=torch.float32, device=xb.device)
        gs = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        wscale = 1.0 / _dist._world(ctx.group) if ctx.group is not None else 1.0
        ex = head._peer_exchange(ctx.group) if ctx.need_dw else None

        def bwd(stage, reserve_sms, ws):
            return ops.head_bwd_impl(xb, w1, b1, w2, b2, targets, seq_off, True, ctx.hp[3], scalars, pre, act, gspill,
                                     row_lse, row_ref, grad_seq, gs, ctx.need_dx, flat, True, ctx.seg[0], ctx.seg[1],
                                     stage, reserve_sms, ws, None, wscale, ex.args if ex is not None else None)

        dx = head._backward_and_sync(bwd, flat, ctx.group, ctx.need_dw, xb, seq_off, ctx.seg)
        gW1 = gB1 = gW2 = gB2 = None
        if ctx.need_dw:
            dW2, dW1, db2, db1 = ops.split_flat_grads(flat, H, E, V)
            d1, d2, d3, d4 = ctx.param_dtypes
            # copies: .grad must never alias the reusable flat buffer
            gW1, gB1, gW2, gB2 = (dW1.to(d1, copy=True), db1.to(d2, copy=True), dW2.to(d3, copy=True),
                                  db2.to(d4, copy=True))
        gx = dx.to(ctx.x_dtype) if ctx.need_dx else None
        return gx, gW1, gB1, gW2, gB2, None, None, None, None, None, None


class _LogpsFn(torch.autograd.Function):
    """seq_logps = get_batch_logps(head(x_rows), targets)  (train.py:357-362), differentiable."""

    @staticmethod
    def forward(ctx, x_rows, W1, B1, W2, B2, head, targets, seq_off, average, group, seg):
        p = head._kernel_params()
        xb = x_rows.detach().to(torch.bfloat16).contiguous()
        need_bwd = any(ctx.needs_input_grad[:5])   # grad mode is off inside forward(); ask autograd instead
        seq_logps, row_logps, row_lse, row_ref, pre, act, gspill = ops.logps_fwd_impl(
            xb, p.w1, p.b1, p.w2, p.b2, targets, seq_off, average, need_bwd, seg[0], seg[1])
        ctx.head, ctx.average, ctx.group, ctx.seg = head, average, group, seg
        ctx.x_dtype = x_rows.dtype
        ctx.need_dx = ctx.needs_input_grad[0]
        ctx.need_dw = any(ctx.needs_input_grad[1:5])
        ctx.param_dtypes = (W1.dtype, B1.dtype, W2.dtype, B2.dtype)
        if need_bwd:
            ctx.save_for_backward(xb, p.w1, p.b1, p.w2, p.b2, targets, seq_off, pre, act, gspill, row_lse, row_ref)
        ctx.mark_non_differentiable(row_logps)
        ctx.set_materialize_grads(False)
        return seq_logps, row_logps

    @staticmethod
    def backward(ctx, grad_seq, _):
        if grad_seq is None:
            return (None,) * 11
        xb, w1, b1, w2, b2, targets, seq_off, pre, act, gspill, row_lse, row_ref = ctx.saved_tensors
        head = ctx.head
        H, E, V = head.n_embed, head.image_token_embed, head.image_token_size
        dev = xb.device
        flat = head._flat_grad_buffer(ctx.group) if ctx.need_dw else torch.empty(0, dtype=torch.float32, device=dev)
        one = torch.ones(1, dtype=torch.float32, device=dev)
        none = torch.empty(0, dtype=torch.float32, device=dev)
        gseq = grad_seq.detach().to(torch.float32).contiguous()
        wscale = 1.0 / _dist._world(ctx.group) if ctx.group is not None else 1.0
        ex = head._peer_exchange(ctx.group) if ctx.need_dw else None

        def bwd(stage, reserve_sms, ws):
            return ops.head_bwd_impl(xb, w1, b1, w2, b2, targets, seq_off, ctx.average, 0.0, none, pre, act, gspill,
                                     row_lse, row_ref, gseq, one, ctx.need_dx, flat, False, ctx.seg[0], ctx.seg[1],
                                     stage, reserve_sms, ws, None, wscale, ex.args if ex is not None else None)

        dx = head._backward_and_sync(bwd, flat, ctx.group, ctx.need_dw, xb, seq_off, ctx.seg)
        gW1 = gB1 = gW2 = gB2 = None
        if ctx.need_dw:
            dW2, dW1, db2, db1 = ops.split_flat_grads(flat, H, E, V)
            d1, d2, d3, d4 = ctx.param_dtypes
            # copies: .grad must never alias the reusable flat buffer
            gW1, gB1, gW2, gB2 = (dW1.to(d1, copy=True), db1.to(d2, copy=True), dW2.to(d3, copy=True),
                                  db2.to(d4, copy=True))
        gx = dx.to(ctx.x_dtype) if ctx.need_dx else None
        return gx, gW1, gB1, gW2, gB2, None, None, None, None, None, None


class FusedGenHead(torch.nn.Module):
    """Linear(n_embed -> image_token_embed) -> exact GELU -> Linear(-> image_token_size), sm_100a only."""

    def __init__(self, params):
        super().__init__()
        self.n_embed = int(params.n_embed)
        self.image_token_embed = int(params.image_token_embed)
        self.image_token_size = int(params.image_token_size)
        # same sub-module names as the reference (modeling_vlm.py:39-45) => identical state-dict keys
        self.output_mlp_projector = torch.nn.Linear(self.n_embed, self.image_token_embed)
        self.vision_activation = torch.nn.GELU()
        self.vision_head = torch.nn.Linear(self.image_token_embed, self.image_token_size)
        self._cache_key = None
        self._cache: Optional[_HeadParams] = None
        self._flat: Optional[torch.Tensor] = None
        self._packed = None
        self._packed_key = None
        self._bwd_count = 0      # fused backward passes since the flat gradient buffer was last consumed / zeroed
        self._flat_is_symmetric = False
        self.decode_packed = True   # decode step streams pre-packed 16 KB weight tiles (False: tensor-map loads of W1 / W2)

    def invalidate(self) -> None:
        """drop the staged kernel operands (bf16 weights, fp32 biases, packed decode weights).  The cache is keyed on
        the parameters' (data_ptr, _version); call this after an update that changes neither -- writes through
        ``.data`` or through views of a flat buffer (apex / DeepSpeed-style optimizers)."""
        self._cache_key = None
        self._cache = None
        self._packed = None
        self._packed_key = None

    def zero_grad(self, set_to_none: bool = True) -> None:
        super().zero_grad(set_to_none=set_to_none)
        self._bwd_count = 0

    # ---- construction helpers -------------------------------------------------------------------
    @classmethod
    def from_reference(cls, module: torch.nn.Module) -> "FusedGenHead":
        """adopt the parameters of a reference ``vision_head`` instance (shares storage, keeps requires_grad)"""
        lin1, lin2 = module.output_mlp_projector, module.vision_head

        class _P:
            n_embed = lin1.in_features
            image_token_embed = lin1.out_features
            image_token_size = lin2.out_features

        new = cls(_P)
        new.output_mlp_projector, new.vision_head = lin1, lin2
        return new

    # ---- parameter staging -------------------------------------------------------------------
    def _kernel_params(self) -> _HeadParams:
        """bf16 weights / fp32 biases as the kernels want them (cached until a parameter changes)"""
        W1, B1 = self.output_mlp_projector.weight, self.output_mlp_projector.bias
        W2, B2 = self.vision_head.weight, self.vision_head.bias
        key = tuple((t.data_ptr(), t._version, t.dtype, t.device) for t in (W1, B1, W2, B2))
        if key != self._cache_key:
            if not W1.is_cuda:
                raise _abi.OspoHeadError("FusedGenHead parameters must live on a B200 (no CPU path): call .cuda()")
            # outside inference mode: a cache first built under @torch.inference_mode() (the reference's generate_image,
            # image_generation.py:109) would hold inference tensors, which a later training call cannot save for backward
            with torch.inference_mode(False), torch.no_grad():
                self._cache = _HeadParams(
                    W1.detach().to(torch.bfloat16).contiguous(), B1.detach().to(torch.float32).contiguous(),
                    W2.detach().to(torch.bfloat16).contiguous(), B2.detach().to(torch.float32).contiguous())
            self._cache_key = key
        return self._cache

    def _decode_packed(self, p: "_HeadParams"):
        """W1 / W2 in the decode kernel's streaming layout (16 KB swizzled tiles), rebuilt when the operands change;
        an extra copy of the weights (168 MB for the 7B head) kept only by heads that generate"""
        key = (p.w1.data_ptr(), p.w2.data_ptr(), self._cache_key)
        if self._packed_key != key or self._packed is None:
            self._packed = (ops.pack_weight_impl(p.w1), ops.pack_weight_impl(p.w2))
            self._packed_key = key
        return self._packed

    def _flat_grad_buffer(self, group=None) -> torch.Tensor:
        n = ops.flat_grad_numel(self.n_embed, self.image_token_embed, self.image_token_size)
        dev = self.vision_head.weight.device
        ex = self._peer_exchange(group)
        if ex is not None:
            self._flat = ex.flat       # symmetric memory: the owners multicast the reduced shards into it
            return self._flat
        if self._flat is None or self._flat.numel() != n or self._flat.device != dev or self._flat_is_symmetric:
            self._flat = torch.empty(n, dtype=torch.float32, device=dev)
            self._flat_is_symmetric = False
        return self._flat

    def _peer_exchange(self, group):
        """the peer-memory gradient exchange for this group (None: single rank, or the NCCL path)"""
        if group is None or _dist._world(group) == 1:
            return None
        ex = _dist.peer_exchange_for(group, self.n_embed, self.image_token_embed, self.image_token_size,
                                     self.vision_head.weight.device)
        if ex is not None:
            self._flat_is_symmetric = True
        return ex

    def _backward_and_sync(self, bwd, flat: torch.Tensor, group, need_dw: bool, xb, seq_off, seg):
        """run the fused backward (``bwd(stage, reserve_sms, workspace) -> dx``) and sum the flat gradient over the
        data-parallel group (the kernels already stored it times 1 / world_size, so the sum is DDP's average).  With
        more than one rank the backward is staged: the exchange of dW2 (80 % of the bytes) runs on a side stream while
        db1 / dW1 / dX are computed.  OSPO_HEAD_OVERLAP=0 restores the single exchange after the backward;
        OSPO_HEAD_DP=p2p selects the NVLink peer-memory exchange instead of the NCCL all-reduce (dist.py).
        (Measured and dropped in round 1: dX as a third part beside the second all-reduce, and SMs reserved for the
        collective.)"""
        world = _dist._world(group) if group is not None else 1
        if need_dw:
            self._bwd_count += 1     # FusedHeadAdamW.step(use_last_backward=True) needs exactly one since the last step
        ex = self._peer_exchange(group) if need_dw else None
        if ex is not None:
            # peer-memory exchange: the backward's weight-gradient stores already went to the owners' inboxes (the
            # reduce-scatter rode inside the GEMM epilogues); what is left is barrier -> each owner sums its inbox and
            # multicasts its shard into everyone's flat buffer -> barrier
            if os.environ.get("OSPO_HEAD_OVERLAP", "1") == "0":
                dx = bwd(0, 0, None)
                ex.finish()
                return dx
            H, E, V = self.n_embed, self.image_token_embed, self.image_token_size
            rows, _ = ops._x_dims(xb, seg[0])
            ws = ops._workspace(rows, H, E, V, seq_off.numel() - 1, xb.device)
            return ex.run_staged(lambda: bwd(1, 0, ws), lambda: bwd(2, 0, ws), lambda: bwd(4, 0, ws))
        if not need_dw or world == 1 or os.environ.get("OSPO_HEAD_OVERLAP", "1") == "0":
            dx = bwd(0, 0, None)
            if need_dw:
                self._sync_flat_grads(flat, group)
            return dx
        H, E, V = self.n_embed, self.image_token_embed, self.image_token_size
        rows, _ = ops._x_dims(xb, seg[0])
        ws = ops._workspace(rows, H, E, V, seq_off.numel() - 1, xb.device)
        return _dist.staged_allreduce_mean_(flat, V * E, group, lambda: bwd(1, 0, ws), lambda: bwd(6, 0, ws),
                                            prescaled=True)

    @staticmethod
    def _sync_flat_grads(flat: torch.Tensor, group) -> None:
        """DDP semantics (ospo/utils/train.py:26-28): average the head-weight gradients over the data-parallel
        ranks -- one NCCL all-reduce (sum; the 1 / world_size factor is already in the buffer) over the contiguous
        fp32 buffer dW2|dW1|db2|db1."""
        if group is None:
            return
        allreduce_mean_(flat, group, prescaled=True)

    # ---- reference-compatible call: logits ---------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``gen_head(hidden_states)`` -> logits [..., V] (bf16), materialised.  Inference path
        (image_generation.py:156); for training use :meth:`logps` / :meth:`simpo`, which never
        materialise the logits and carry the backward."""
        if torch.is_grad_enabled() and (x.requires_grad or self.vision_head.weight.requires_grad):
            raise _abi.OspoHeadError(
                "FusedGenHead.forward materialises logits for inference only; call it under torch.no_grad() or use "
                ".logps()/.simpo() for the differentiable fused path")
        p = self._kernel_params()
        lead = x.shape[:-1]
        xb = x.reshape(-1, self.n_embed).to(torch.bfloat16).contiguous()
        out = ops.linear_gelu_linear_impl(xb, p.w1, p.b1, p.w2, p.b2)
        return out.view(*lead, self.image_token_size)

    # ---- fused training paths ----------------------------------------------------------------
    def logps(self, hidden: torch.Tensor, labels: torch.Tensor, average_log_prob: bool = True,
              ignore_index: int = IGNORE_INDEX, image_span: Optional[Tuple[int, int]] = None,
              process_group=None, return_per_token: bool = False):
        """== ``get_batch_logps(gen_head(hidden), labels, average_log_prob)`` (train.py:357-362, 375-396)
        hidden [S, L, H], labels [S, L] (unshifted, ``ignore_index`` on masked positions) -> [S] fp32."""
        x_rows, targets, seq_off, seg = _rows_from_labels(hidden, labels, ignore_index, image_span)
        seq_logps, row_logps = _LogpsFn.apply(
            x_rows, self.output_mlp_projector.weight, self.output_mlp_projector.bias, self.vision_head.weight,
            self.vision_head.bias, self, targets, seq_off, bool(average_log_prob), process_group, seg)
        return (seq_logps, row_logps) if return_per_token else seq_logps

    def simpo(self, hidden: torch.Tensor, labels: torch.Tensor, *, beta: float = 1.0, gamma_beta_ratio: float = 0.0,
              label_smoothing: float = 0.0, loss_type: str = "sigmoid", sft_weight: float = 0.0,
              ignore_index: int = IGNORE_INDEX, image_span: Optional[Tuple[int, int]] = None,
              process_group=None) -> SimpoOutput:
        """Fused ``concatenated_forward`` + ``simpo_loss`` + ``losses.mean()`` (train.py:345-372, 317-342, 419-430).
        hidden [2B, L, H] = chosen sequences then rejected sequences (train.py:364-365), labels [2B, L]."""
        if loss_type not in ("sigmoid", "hinge"):
            raise ValueError(f"Unknown loss type: {loss_type}. Should be one of ['sigmoid', 'hinge']")  # train.py:336
        lt = _abi.LOSS_SIGMOID if loss_type == "sigmoid" else _abi.LOSS_HINGE
        x_rows, targets, seq_off, seg = _rows_from_labels(hidden, labels, ignore_index, image_span)
        hp = (float(beta), float(gamma_beta_ratio), float(label_smoothing), float(sft_weight), lt)
        loss, scalars, seq_logps, losses, crew, rrew, row_logps = _SimpoFn.apply(
            x_rows, self.output_mlp_projector.weight, self.output_mlp_projector.bias, self.vision_head.weight,
            self.vision_head.bias, self, targets, seq_off, hp, process_group, seg)
        B = seq_logps.shape[0] // 2
        metrics = {
            "rewards/chosen": scalars[_abi.SC_REWARD_CHOSEN], "rewards/rejected": scalars[_abi.SC_REWARD_REJECTED],
            "rewards/accuracies": scalars[_abi.SC_REWARD_ACC], "rewards/margins": scalars[_abi.SC_REWARD_MARGIN],
            "logps/chosen": scalars[_abi.SC_LOGPS_CHOSEN], "logps/rejected": scalars[_abi.SC_LOGPS_REJECTED],
            "logits/chosen": scalars[_abi.SC_LOGITS_CHOSEN], "logits/rejected": scalars[_abi.SC_LOGITS_REJECTED],
            "sft_loss": scalars[_abi.SC_SFT_LOSS], "simpo_loss": scalars[_abi.SC_SIMPO_LOSS],
        }
        return SimpoOutput(loss, seq_logps[:B], seq_logps[B:], losses, crew, rrew, row_logps, metrics)

    # ---- CFG decode step ---------------------------------------------------------------------
    @torch.no_grad()
    def cfg_sample(self, hidden_last: torch.Tensor, cfg_weight: float = 5.0, temperature: float = 1.0,
                   uniforms: Optional[torch.Tensor] = None, greedy: bool = False, merge_mode: str = "bf16",
                   return_logits: bool = False, out: Optional[torch.Tensor] = None,
                   next_embeds: Optional["FusedGenImgEmbeds"] = None, embeds_out: Optional[torch.Tensor] = None):
        """One decode step (image_generation.py:156-164): hidden_last [2P, H] with row 2k conditional and
        2k+1 unconditional -> next_token ids [P] int64.  ``uniforms`` [P] fp32 in [0,1) drive the inverse-CDF
        draw (None => drawn from torch's CUDA generator); ``greedy`` takes the arg-max instead; ``out`` (int64
        [P], e.g. a row of the generated-token buffer) receives the ids without a copy.  With ``next_embeds`` (a
        ``FusedGenImgEmbeds``) the call also returns the next step's input embeddings [2P, D]
        (image_generation.py:166-168), produced in the same launch chain: ``(ids, embeds)``."""
        p = self._kernel_params()
        h = hidden_last.to(torch.bfloat16).contiguous()
        P = h.shape[0] // 2
        if greedy:
            u = torch.empty(0, dtype=torch.float32, device=h.device)
        elif uniforms is None:
            u = torch.rand(P, dtype=torch.float32, device=h.device)
        else:
            u = uniforms.to(torch.float32).contiguous()
        mm = _abi.MERGE_BF16 if merge_mode == "bf16" else _abi.MERGE_FP32
        ne = None
        if next_embeds is not None:
            e, wa, ba, wb, bb = next_embeds._params()
            if embeds_out is None:
                embeds_out = torch.empty(2 * P, wb.shape[0], dtype=torch.bfloat16, device=h.device)
            ne = (e, wa, ba, wb, bb, embeds_out, next_embeds._table_or_none())
        packed = self._decode_packed(p) if (h.shape[0] <= 32 and self.decode_packed) else None
        ids, logits = ops.cfg_sample_impl(h, p.w1, p.b1, p.w2, p.b2, float(cfg_weight), float(temperature), u,
                                          bool(greedy), mm, bool(return_logits), out, ne, packed)
        if next_embeds is not None:
            return (ids, logits, embeds_out) if return_logits else (ids, embeds_out)
        return (ids, logits) if return_logits else ids


@torch.no_grad()
def cfg_merge_sample(logits: torch.Tensor, cfg_weight: float = 5.0, temperature: float = 1.0,
                     uniforms: Optional[torch.Tensor] = None, greedy: bool = False, merge_mode: str = "bf16",
                     return_merged: bool = False):
    """image_generation.py:157-163 on supplied bf16 logits [..., 2P, V] -> ids [..., P]."""
    lg = logits.to(torch.bfloat16).contiguous()
    P = lg.shape[-2] // 2
    n = lg.numel() // (lg.shape[-1] * lg.shape[-2]) * P
    if greedy:
        u = torch.empty(0, dtype=torch.float32, device=lg.device)
    elif uniforms is None:
        u = torch.rand(n, dtype=torch.float32, device=lg.device)
    else:
        u = uniforms.to(torch.float32).contiguous()
    mm = _abi.MERGE_BF16 if merge_mode == "bf16" else _abi.MERGE_FP32
    ids, merged = ops.cfg_merge_sample_impl(lg, float(cfg_weight), float(temperature), u, bool(greedy), mm,
                                            bool(return_merged))
    return (ids, merged) if return_merged else ids


class FusedGenImgEmbeds:
    """Drop-in for ``MultiModalityCausalLM.prepare_gen_img_embeds`` (janus/models/modeling_vlm.py:263-264):
    ``gen_aligner(gen_embed(image_ids))`` with the reference's modules -- ``gen_embed = nn.Embedding(16384, 8)``,
    ``gen_aligner = MlpProjector('mlp_gelu', depth=2)`` i.e. ``layers = Sequential(Linear(8, D), GELU(), Linear(D, D))``
    (janus/models/projector.py:39-45).  Holds no parameters of its own: it reads the modules' tensors (bf16 / fp32
    staging cached until they change).  Inference only, like the call site (image_generation.py:166-168)."""

    def __init__(self, gen_embed: torch.nn.Embedding, gen_aligner: torch.nn.Module):
        layers = gen_aligner.layers
        if not (isinstance(layers, torch.nn.Sequential) and len(layers) == 3 and isinstance(layers[0], torch.nn.Linear)
                and isinstance(layers[2], torch.nn.Linear) and isinstance(layers[1], torch.nn.GELU)):
            raise _abi.OspoHeadError("FusedGenImgEmbeds supports the 'mlp_gelu' depth-2 aligner of Janus-Pro only")
        if gen_embed.embedding_dim != 8 or layers[0].in_features != 8:
            raise _abi.OspoHeadError("FusedGenImgEmbeds expects the 8-dimensional VQ code embedding")
        self.gen_embed, self.lin_a, self.lin_b = gen_embed, layers[0], layers[2]
        self._key, self._staged = None, None
        self._table, self._table_key = None, None
        self.use_table = False

    def invalidate(self) -> None:
        """drop the staged operands and the memo table (see FusedGenHead.invalidate)"""
        self._key, self._staged = None, None
        self._table, self._table_key = None, None

    def build_table(self) -> torch.Tensor:
        """Memoise the module over the whole VQ codebook: ``table[id] = gen_aligner(gen_embed(id))`` for every id
        (bf16 [codebook, D]; 134 MB for Janus-Pro-7B, built in a few ms with the same kernels, so every row holds
        exactly the bits a per-step evaluation gives).  ``gen_aligner(gen_embed(.))`` is a pure function of the token
        id and generation runs with frozen weights (image_generation.py:109 is under inference_mode), so the decode
        loop can then fetch the next step's embeddings as two rows of this table: no 33.6 MB weight stream and no
        extra kernel per step.  Rebuilt automatically when a parameter changes; ``use_table = False`` switches back."""
        staged = self._params()
        if self._table is None or self._table_key != self._key:
            e = staged[0]
            ids = torch.arange(e.shape[0], dtype=torch.int64, device=e.device)
            with torch.inference_mode(False), torch.no_grad():
                self._table = ops.gen_img_embeds_impl(ids, *staged)
            self._table_key = self._key
        self.use_table = True
        return self._table

    def _table_or_none(self):
        if not self.use_table:
            return None
        return self.build_table()

    def _params(self):
        ts = (self.gen_embed.weight, self.lin_a.weight, self.lin_a.bias, self.lin_b.weight, self.lin_b.bias)
        key = tuple((t.data_ptr(), t._version, t.dtype, t.device) for t in ts)
        if key != self._key:
            with torch.inference_mode(False), torch.no_grad():
                e, wa, ba, wb, bb = ts
                self._staged = (e.detach().to(torch.bfloat16).contiguous(), wa.detach().to(torch.bfloat16).contiguous(),
                                ba.detach().to(torch.float32).contiguous(), wb.detach().to(torch.bfloat16).contiguous(),
                                bb.detach().to(torch.float32).contiguous())
            self._key = key
        return self._staged

    @torch.no_grad()
    def __call__(self, image_ids: torch.Tensor) -> torch.Tensor:
        e, wa, ba, wb, bb = self._params()
        ids = image_ids.reshape(-1).to(torch.int64).contiguous()
        out = ops.gen_img_embeds_impl(ids, e, wa, ba, wb, bb, table=self._table_or_none())
        return out.view(*image_ids.shape, out.shape[-1])

    @torch.no_grad()
    def from_sampled(self, next_token: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """image_generation.py:166-168 in one call: ``next_token`` [P] (the sampler's ids) -> the next step's
        ``inputs_embeds`` rows [2P, D] (row 2k and 2k+1 both come from id k, the cond/uncond duplication), written
        into ``out`` if given.  No ``torch.cat`` / ``view`` / copy kernels: the launch chain stays dependent-launch
        linked to the sampler."""
        e, wa, ba, wb, bb = self._params()
        ids = next_token.reshape(-1).to(torch.int64).contiguous()
        return ops.gen_img_embeds_impl(ids, e, wa, ba, wb, bb, 2, out, self._table_or_none())
