"""torch custom ops (``torch.ops.ospo_head.*``) over the C ABI.

Each op takes/returns plain tensors, allocates its outputs and scratch with torch (the library itself
never allocates device memory) and enqueues the kernels on torch's current CUDA stream.  Only a CUDA
implementation is registered: calling an op with CPU tensors fails in the dispatcher, and a missing
``libospo_head.so`` raises at first use -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _abi

Tensor = torch.Tensor


def _ptr(t: Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _abi.OspoHeadError("ospo_head ops run on a B200 only: got a CPU tensor (there is no CPU path)")


def _weights(w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> _abi.Weights:
    assert w1.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16, "weights must be bf16"
    assert b1.dtype == torch.float32 and b2.dtype == torch.float32, "biases must be fp32"
    assert w1.is_contiguous() and w2.is_contiguous() and b1.is_contiguous() and b2.is_contiguous()
    return _abi.Weights(w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr())


def _workspace(rows: int, H: int, E: int, V: int, S: int, device) -> Tensor:
    n = _abi.workspace_bytes(rows, H, E, V, S)
    return torch.empty(n, dtype=torch.uint8, device=device)


def flat_grad_numel(H: int, E: int, V: int) -> int:
    return V * E + E * H + V + E


def split_flat_grads(flat: Tensor, H: int, E: int, V: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """views (dW2 [V,E], dW1 [E,H], db2 [V], db1 [E]) of the flat gradient buffer"""
    o0, o1, o2 = V * E, V * E + E * H, V * E + E * H + V
    return flat[:o0].view(V, E), flat[o0:o1].view(E, H), flat[o1:o2], flat[o2:o2 + E]


# --------------------------------------------------------------------------------------------------
# plain logits
# --------------------------------------------------------------------------------------------------
def linear_gelu_linear_impl(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
    """logits[rows, V] (bf16) = vision_head(x[rows, H]);  janus/models/modeling_vlm.py:47-51"""
    _check_cuda(x, w1, b1, w2, b2)
    assert x.dim() == 2 and x.dtype == torch.bfloat16 and x.is_contiguous()
    rows, H = x.shape
    E, V = w1.shape[0], w2.shape[0]
    logits = torch.empty(rows, V, dtype=torch.bfloat16, device=x.device)
    ws = _workspace(rows, H, E, V, 1, x.device)
    args = _abi.HeadArgs(_abi.Shape(rows, H, E, V, 1), _weights(w1, b1, w2, b2), x.data_ptr(), logits.data_ptr(),
                         ws.data_ptr(), ws.numel())
    _abi.check(_abi.load().ospo_head_logits(C.byref(args), _stream()), "ospo_head_logits")
    return logits




# --------------------------------------------------------------------------------------------------
# log-probs / SimPO forward.  Returns the tensors the backward needs as explicit outputs.
# --------------------------------------------------------------------------------------------------
def _x_dims(x: Tensor, seg_rows: int):
    """(rows, H) of the head's input: x is [rows, H], or -- row-segmented -- [S, pitch, H] with seg_rows rows used"""
    if seg_rows:
        return x.shape[0] * seg_rows, x.shape[2]
    return x.shape[0], x.shape[1]


def _simpo_args(x, w1, b1, w2, b2, labels, seq_off, average, hp, outs, saved, bwd, ws, seg=(0, 0, 0)) -> _abi.SimpoArgs:
    rows, H = _x_dims(x, seg[0])
    E, V = w1.shape[0], w2.shape[0]
    S = seq_off.numel() - 1
    a = _abi.SimpoArgs()
    a.shape = _abi.Shape(rows, H, E, V, S)
    a.w = _weights(w1, b1, w2, b2)
    a.x = x.data_ptr()
    a.x_seg_rows, a.x_seg_pitch, a.x_seg_off = seg
    a.labels = labels.data_ptr()
    a.seq_offsets = seq_off.data_ptr()
    a.average_log_prob = int(average)
    a.beta, a.gamma_beta_ratio, a.label_smoothing, a.sft_weight, a.loss_type = hp
    (a.row_logps, a.seq_logps, a.losses, a.chosen_rewards, a.rejected_rewards, a.scalars) = [_ptr(t) for t in outs]
    (a.pre, a.act, a.logits, a.row_lse, a.row_ref, a.grad_seq) = [_ptr(t) for t in saved]
    (a.grad_loss, a.dx, a.flat_grads) = [_ptr(t) for t in bwd]
    a.workspace = ws.data_ptr()
    a.workspace_bytes = ws.numel()
    return a


def _check_rows(x: Tensor, labels: Tensor, seq_off: Tensor, seg_rows: int = 0) -> None:
    _check_cuda(x, labels, seq_off)
    assert x.dtype == torch.bfloat16 and x.is_contiguous(), "x must be contiguous bf16"
    assert x.dim() == (3 if seg_rows else 2), "x is [rows, H], or [S, pitch, H] when row-segmented"
    rows, _ = _x_dims(x, seg_rows)
    assert labels.dtype == torch.int64 and labels.shape == (rows,) and labels.is_contiguous()
    assert seq_off.dtype == torch.int64 and seq_off.dim() == 1 and seq_off.is_contiguous()


def logps_fwd_impl(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, labels: Tensor, seq_off: Tensor,
              average: bool, save_for_backward: bool, seg_rows: int = 0, seg_off: int = 0) -> List[Tensor]:
    """-> [seq_logps[S], row_logps[rows], row_lse[rows], row_ref[rows], pre, act, gspill]  (train.py:357 + 375-396)
    seg_rows > 0: x is the [S, pitch, H] hidden-state tensor and rows [seg_off, seg_off + seg_rows) of every
    sequence are the head's rows (no gather copy).  ``gspill`` [rows, V] bf16 is the forward's only logits-sized
    product: the unscaled softmax-minus-onehot numerator the backward GEMMs read (never modified afterwards)."""
    _check_rows(x, labels, seq_off, seg_rows)
    rows, H = _x_dims(x, seg_rows)
    seg = (seg_rows, x.shape[1], seg_off) if seg_rows else (0, 0, 0)
    E, V = w1.shape[0], w2.shape[0]
    S = seq_off.numel() - 1
    dev = x.device
    f32 = dict(dtype=torch.float32, device=dev)
    seq_logps, row_logps, row_lse = torch.empty(S, **f32), torch.empty(rows, **f32), torch.empty(rows, **f32)
    row_ref = torch.empty(rows, **f32)
    act = torch.empty(rows, E, dtype=torch.bfloat16, device=dev)
    if save_for_backward:
        pre = torch.empty(rows, E, dtype=torch.bfloat16, device=dev)
        logits = torch.empty(rows, V, dtype=torch.bfloat16, device=dev)
    else:
        pre = logits = None
    ws = _workspace(rows, H, E, V, S, dev)
    a = _simpo_args(x, w1, b1, w2, b2, labels, seq_off, average, (1.0, 0.0, 0.0, 0.0, 0),
                    (row_logps, seq_logps, None, None, None, None), (pre, act, logits, row_lse, row_ref, None),
                    (None, None, None), ws, seg)
    _abi.check(_abi.load().ospo_head_logps_fwd(C.byref(a), _stream()), "ospo_head_logps_fwd")
    def empty():
        return torch.empty(0, dtype=torch.bfloat16, device=dev)

    return [seq_logps, row_logps, row_lse, row_ref, pre if pre is not None else empty(), act,
            logits if logits is not None else empty()]


def simpo_fwd_impl(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, labels: Tensor, seq_off: Tensor,
              beta: float, gamma_beta_ratio: float, label_smoothing: float, sft_weight: float, loss_type: int,
              save_for_backward: bool, seg_rows: int = 0, seg_off: int = 0) -> List[Tensor]:
    """-> [scalars[16], seq_logps[S], losses[B], chosen_rewards[B], rejected_rewards[B], row_logps[rows],
           row_lse[rows], row_ref[rows], grad_seq[S], pre, act, gspill]      (train.py:345-372, 317-342, 399-445)"""
    _check_rows(x, labels, seq_off, seg_rows)
    rows, H = _x_dims(x, seg_rows)
    seg = (seg_rows, x.shape[1], seg_off) if seg_rows else (0, 0, 0)
    E, V = w1.shape[0], w2.shape[0]
    S = seq_off.numel() - 1
    assert S % 2 == 0, "SimPO needs chosen and rejected halves"
    B = S // 2
    dev = x.device
    f32 = dict(dtype=torch.float32, device=dev)
    scalars = torch.zeros(_abi.SC_COUNT, **f32)
    seq_logps, losses = torch.empty(S, **f32), torch.empty(B, **f32)
    crew, rrew = torch.empty(B, **f32), torch.empty(B, **f32)
    row_logps, row_lse, grad_seq = torch.empty(rows, **f32), torch.empty(rows, **f32), torch.empty(S, **f32)
    row_ref = torch.empty(rows, **f32)
    act = torch.empty(rows, E, dtype=torch.bfloat16, device=dev)
    if save_for_backward:
        pre = torch.empty(rows, E, dtype=torch.bfloat16, device=dev)
        logits = torch.empty(rows, V, dtype=torch.bfloat16, device=dev)
    else:
        pre = logits = None
    ws = _workspace(rows, H, E, V, S, dev)
    a = _simpo_args(x, w1, b1, w2, b2, labels, seq_off, True,
                    (beta, gamma_beta_ratio, label_smoothing, sft_weight, loss_type),
                    (row_logps, seq_logps, losses, crew, rrew, scalars), (pre, act, logits, row_lse, row_ref, grad_seq),
                    (None, None, None), ws, seg)
    _abi.check(_abi.load().ospo_head_simpo_fwd(C.byref(a), _stream()), "ospo_head_simpo_fwd")
    def empty():
        return torch.empty(0, dtype=torch.bfloat16, device=dev)

    return [scalars, seq_logps, losses, crew, rrew, row_logps, row_lse, row_ref, grad_seq,
            pre if pre is not None else empty(), act, logits if logits is not None else empty()]


def head_bwd_impl(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, labels: Tensor, seq_off: Tensor,
             average: bool, sft_weight: float, scalars: Tensor, pre: Tensor, act: Tensor, logits: Tensor,
             row_lse: Tensor, row_ref: Tensor, grad_seq: Tensor, grad_scale: Tensor, need_dx: bool, flat_grads: Tensor,
             simpo: bool, seg_rows: int = 0, seg_off: int = 0, stage: int = 0, reserve_sms: int = 0,
             ws: Optional[Tensor] = None, dx_out: Optional[Tensor] = None, wgrad_scale: float = 1.0,
             dp=None) -> Tensor:
    """the dgrad / wgrad GEMM pairs on the forward's softmax-minus-onehot spill.
    ``stage`` is a bit mask (0 = everything): 1 = up to dW2, 2 = db1 + dW1, 4 = dX, so the caller can overlap the
    all-reduces with the later parts (pass the same ``ws`` to every call; dx is produced by part 4).
    `logits` (the spill) is only read; `flat_grads` (numel 0 = head frozen) receives wgrad_scale * (dW2|dW1|db2|db1).
    Returns dx bf16 with the shape of x (numel 0 if not requested); with a row-segmented x the rows outside the
    span are zero (train.py: masked positions carry no gradient)."""
    _check_rows(x, labels, seq_off, seg_rows)
    rows, H = _x_dims(x, seg_rows)
    E, V = w1.shape[0], w2.shape[0]
    S = seq_off.numel() - 1
    dev = x.device
    seg = (seg_rows, x.shape[1], seg_off) if seg_rows else (0, 0, 0)
    dx = None
    if need_dx and (stage == 0 or (stage & 4)):
        dx = torch.empty_like(x) if dx_out is None else dx_out
        if seg_rows:
            dx[:, :seg_off].zero_()
            dx[:, seg_off + seg_rows:].zero_()
    fg = flat_grads if flat_grads.numel() else None
    if fg is not None:
        assert fg.dtype == torch.float32 and fg.numel() == flat_grad_numel(H, E, V) and fg.is_contiguous()
    assert grad_scale.dtype == torch.float32 and grad_scale.numel() == 1
    assert grad_seq.dtype == torch.float32 and grad_seq.numel() == S and grad_seq.is_contiguous()
    if ws is None:
        ws = _workspace(rows, H, E, V, S, dev)
    a = _simpo_args(x, w1, b1, w2, b2, labels, seq_off, average, (1.0, 0.0, 0.0, sft_weight, 0),
                    (None, None, None, None, None, scalars if scalars.numel() else None),
                    (pre, act, logits, row_lse, row_ref, grad_seq), (grad_scale, dx, fg), ws, seg)
    a.bwd_stage, a.reserve_sms = int(stage), int(reserve_sms)
    a.wgrad_scale = float(wgrad_scale)
    if dp is not None:          # _abi.DpExchange: the weight-gradient stores go to the owners' inboxes (peer memory)
        a.dp = C.cast(C.pointer(dp), C.c_void_p)
    lib = _abi.load()
    if simpo:
        _abi.check(lib.ospo_head_simpo_bwd(C.byref(a), _stream()), "ospo_head_simpo_bwd")
    else:
        _abi.check(lib.ospo_head_logps_bwd(C.byref(a), _stream()), "ospo_head_logps_bwd")
    return dx if dx is not None else torch.empty(0, dtype=torch.bfloat16, device=dev)


# --------------------------------------------------------------------------------------------------
# CFG decode step
# --------------------------------------------------------------------------------------------------
def pack_weight_impl(w: Tensor) -> Tensor:
    """[rows, cols] bf16 row-major weight -> the decode kernel's streaming layout (16 KB swizzled tiles, uint8 buffer)"""
    _check_cuda(w)
    assert w.dtype == torch.bfloat16 and w.dim() == 2 and w.is_contiguous()
    need = C.c_size_t()
    lib = _abi.load()
    _abi.check(lib.ospo_head_packed_weight_bytes(w.shape[0], w.shape[1], C.byref(need)), "ospo_head_packed_weight_bytes")
    out = torch.empty(need.value, dtype=torch.uint8, device=w.device)
    _abi.check(lib.ospo_head_pack_weight(w.data_ptr(), w.shape[0], w.shape[1], out.data_ptr(), _stream()),
               "ospo_head_pack_weight")
    return out


def cfg_sample_impl(h: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, cfg_weight: float, temperature: float,
               uniforms: Tensor, greedy: bool, merge_mode: int, want_logits: bool = False,
               out: Optional[Tensor] = None, next_embeds: Optional[tuple] = None,
               packed: Optional[tuple] = None) -> List[Tensor]:
    """-> [ids[P] int64 (``out`` if given: the kernel writes the ids there), logits[2P, V] bf16 (empty unless want_logits)]
    (image_generation.py:156-164; row 2k cond / 2k+1 uncond).  Without want_logits the logits never leave the
    chip: the CFG merge, softmax weights and segment sums are produced in the GEMM epilogue."""
    _check_cuda(h, uniforms)
    assert h.dim() == 2 and h.dtype == torch.bfloat16 and h.is_contiguous() and h.shape[0] % 2 == 0
    rows, H = h.shape
    E, V = w1.shape[0], w2.shape[0]
    P = rows // 2
    dev = h.device
    if not greedy:
        assert uniforms.dtype == torch.float32 and uniforms.numel() == P and uniforms.is_contiguous()
    if out is not None:
        assert out.dtype == torch.int64 and out.numel() == P and out.is_contiguous() and out.device == dev
        ids = out
    else:
        ids = torch.empty(P, dtype=torch.int64, device=dev)
    logits = torch.empty(rows, V, dtype=torch.bfloat16, device=dev) if want_logits else None
    ws = _workspace(rows, H, E, V, 1, dev)
    a = _abi.CfgArgs()
    a.shape = _abi.Shape(rows, H, E, V, 1)
    a.w = _weights(w1, b1, w2, b2)
    a.h = h.data_ptr()
    a.logits = _ptr(logits)
    a.cfg_weight, a.temperature, a.merge_mode, a.greedy, a.num_steps = cfg_weight, temperature, merge_mode, int(greedy), 1
    a.uniforms = None if greedy else uniforms.data_ptr()
    a.ids = ids.data_ptr()
    a.merged = None
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    if packed is not None:                      # (pack_weight_impl(w1), pack_weight_impl(w2))
        a.w1_packed, a.w2_packed = packed[0].data_ptr(), packed[1].data_ptr()
    if next_embeds is not None:
        # (gen_embed, wa, ba, wb, bb, embeds_out[2P, D][, table]): the next step's input embeddings,
        # image_generation.py:166-168, produced in the same launch chain (first aligner layer inside the sampler's
        # finish kernel, then the D x D Linear; with the memo table [codebook, D] the finish kernel copies the rows
        # and nothing else runs)
        ge, wa, ba, wb, bb, eo = next_embeds[:6]
        table = next_embeds[6] if len(next_embeds) > 6 else None
        D = wb.shape[0]
        assert rows <= 32 and eo.dtype == torch.bfloat16 and eo.is_contiguous() and eo.numel() == rows * D
        if table is not None:
            assert table.dtype == torch.bfloat16 and table.is_contiguous() and table.shape == (ge.shape[0], D)
            al = _abi.AlignerArgs(rows, D, ge.shape[0], 8, None, None, None, None, None, None, eo.data_ptr(), None, 0, 2,
                                  table.data_ptr())
        else:
            ws2 = torch.empty(rows * D, dtype=torch.bfloat16, device=dev)
            al = _abi.AlignerArgs(rows, D, ge.shape[0], 8, None, ge.data_ptr(), wa.data_ptr(), ba.data_ptr(),
                                  wb.data_ptr(), bb.data_ptr(), eo.data_ptr(), ws2.data_ptr(), ws2.numel() * 2, 2, None)
        a.next_embeds = C.cast(C.pointer(al), C.c_void_p)
    _abi.check(_abi.load().ospo_head_cfg_sample(C.byref(a), _stream()), "ospo_head_cfg_sample")
    return [ids, logits if logits is not None else torch.empty(0, dtype=torch.bfloat16, device=dev)]


def cfg_merge_sample_impl(logits: Tensor, cfg_weight: float, temperature: float, uniforms: Tensor, greedy: bool,
                     merge_mode: int, want_merged: bool) -> List[Tensor]:
    """merge + softmax + sample on supplied bf16 logits [..., 2P, V] (any number of leading "step" dims)
    -> [ids[..., P] int64, merged[..., P, V] fp32 or empty]"""
    _check_cuda(logits, uniforms)
    assert logits.dtype == torch.bfloat16 and logits.is_contiguous() and logits.dim() >= 2
    V = logits.shape[-1]
    rows = logits.shape[-2]
    steps = logits.numel() // (rows * V)
    P = rows // 2
    dev = logits.device
    if not greedy:
        assert uniforms.dtype == torch.float32 and uniforms.numel() == steps * P and uniforms.is_contiguous()
    ids = torch.empty(*logits.shape[:-2], P, dtype=torch.int64, device=dev)
    merged = torch.empty(*logits.shape[:-2], P, V, dtype=torch.float32, device=dev) if want_merged else None
    a = _abi.CfgArgs()
    a.shape = _abi.Shape(rows, 8, 8, V, 1)
    a.logits = logits.data_ptr()
    a.cfg_weight, a.temperature, a.merge_mode, a.greedy, a.num_steps = cfg_weight, temperature, merge_mode, int(greedy), steps
    a.uniforms = None if greedy else uniforms.data_ptr()
    a.ids = ids.data_ptr()
    a.merged = _ptr(merged)
    _abi.check(_abi.load().ospo_head_cfg_merge_sample(C.byref(a), _stream()), "ospo_head_cfg_merge_sample")
    return [ids, merged if merged is not None else torch.empty(0, dtype=torch.float32, device=dev)]


# --------------------------------------------------------------------------------------------------
# next row N1: sampled ids -> next-step input embeddings
# --------------------------------------------------------------------------------------------------
def gen_img_embeds_impl(ids: Tensor, gen_embed: Tensor, wa: Tensor, ba: Tensor, wb: Tensor, bb: Tensor,
                        repeat: int = 1, out: Optional[Tensor] = None, table: Optional[Tensor] = None) -> Tensor:
    """== gen_aligner(gen_embed(ids))  (janus/models/modeling_vlm.py:263-264; MlpProjector mlp_gelu depth 2)
    ids [n] int64 -> bf16 [n * repeat, D]; row i*repeat + j comes from ids[i] (repeat = 2 is the cond/uncond
    duplication of image_generation.py:166).  At most 32 output rows per launch; longer inputs go in slices.
    ``out`` (bf16 [n * repeat, D], contiguous) receives the result without a copy.  ``table`` (bf16 [codebook, D], the
    module evaluated once for every code) turns the call into a row gather."""
    _check_cuda(ids, gen_embed, wa, wb)
    assert ids.dtype == torch.int64 and ids.dim() == 1 and ids.is_contiguous()
    assert gen_embed.dtype == torch.bfloat16 and gen_embed.shape[1] == 8 and gen_embed.is_contiguous()
    assert wa.dtype == torch.bfloat16 and wb.dtype == torch.bfloat16 and wa.is_contiguous() and wb.is_contiguous()
    assert ba.dtype == torch.float32 and bb.dtype == torch.float32
    D = wb.shape[0]
    n = ids.numel()
    rep = max(1, int(repeat))
    assert 32 % rep == 0, "repeat must divide 32"
    if out is None:
        out = torch.empty(n * rep, D, dtype=torch.bfloat16, device=ids.device)
    else:
        assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() == n * rep * D and out.device == ids.device
    out2 = out.view(n * rep, D)
    ws = torch.empty(32 * D, dtype=torch.bfloat16, device=ids.device)
    lib = _abi.load()
    per = 32 // rep
    for lo in range(0, n, per):
        m = min(per, n - lo)
        a = _abi.AlignerArgs(m * rep, D, gen_embed.shape[0], 8, ids[lo:lo + m].data_ptr(), gen_embed.data_ptr(),
                             wa.data_ptr(), ba.data_ptr(), wb.data_ptr(), bb.data_ptr(),
                             out2[lo * rep:(lo + m) * rep].data_ptr(), ws.data_ptr(), ws.numel() * 2, rep,
                             None if table is None else table.data_ptr())
        _abi.check(lib.ospo_head_gen_img_embeds(C.byref(a), _stream()), "ospo_head_gen_img_embeds")
    return out


# --------------------------------------------------------------------------------------------------
# next row N3: gradient-norm clip + AdamW on the flat buffer
# --------------------------------------------------------------------------------------------------
def grad_sqnorm_impl(flat_grads: Tensor) -> Tensor:
    """device scalar = sum(flat_grads ** 2), summed in a fixed order (the head's share of clip_grad_norm_'s norm)"""
    _check_cuda(flat_grads)
    assert flat_grads.dtype == torch.float32 and flat_grads.is_contiguous()
    out = torch.empty(1, dtype=torch.float32, device=flat_grads.device)
    ws = torch.empty(4096, dtype=torch.float32, device=flat_grads.device)
    _abi.check(_abi.load().ospo_head_grad_sqnorm(flat_grads.data_ptr(), flat_grads.numel(), out.data_ptr(), ws.data_ptr(),
                                                 ws.numel() * 4, _stream()), "ospo_head_grad_sqnorm")
    return out


def adamw_step_impl(flat_grads: Tensor, params: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step: int, lr: float,
                    beta1: float, beta2: float, eps: float, weight_decay: float, max_norm: float = 0.0,
                    total_sqnorm: Optional[Tensor] = None, params_bf16: Optional[Tensor] = None) -> None:
    """clip (coef from ``total_sqnorm``) + torch.optim.AdamW update, in place on params / exp_avg / exp_avg_sq (fp32,
    layout of the flat gradient); ``params_bf16`` (a prefix-sized bf16 buffer) receives the refreshed GEMM operands"""
    _check_cuda(flat_grads, params, exp_avg, exp_avg_sq)
    n = flat_grads.numel()
    for t in (flat_grads, params, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    a = _abi.AdamWArgs()
    a.numel = n
    a.grads, a.params, a.exp_avg, a.exp_avg_sq = flat_grads.data_ptr(), params.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr()
    if params_bf16 is not None:
        assert params_bf16.dtype == torch.bfloat16 and params_bf16.is_contiguous() and params_bf16.numel() <= n
        a.params_bf16, a.shadow_numel = params_bf16.data_ptr(), params_bf16.numel()
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = lr, beta1, beta2, eps, weight_decay
    a.step, a.max_norm = int(step), float(max_norm)
    if max_norm > 0:
        assert total_sqnorm is not None and total_sqnorm.dtype == torch.float32 and total_sqnorm.is_cuda
        a.total_sqnorm = total_sqnorm.data_ptr()
    _abi.check(_abi.load().ospo_head_adamw_step(C.byref(a), _stream()), "ospo_head_adamw_step")


# --------------------------------------------------------------------------------------------------
# registration: torch.ops.ospo_head.<name>  (CUDA only).  The nn.Module in head.py calls the *_impl
# functions directly to keep the dispatcher out of the 576-step decode loop.
# --------------------------------------------------------------------------------------------------
linear_gelu_linear = torch.library.custom_op("ospo_head::linear_gelu_linear", linear_gelu_linear_impl,
                                             mutates_args=(), device_types="cuda")
logps_fwd = torch.library.custom_op("ospo_head::logps_fwd", logps_fwd_impl, mutates_args=(), device_types="cuda")
simpo_fwd = torch.library.custom_op("ospo_head::simpo_fwd", simpo_fwd_impl, mutates_args=(), device_types="cuda")
def _head_bwd_op(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, labels: Tensor, seq_off: Tensor,
                 average: bool, sft_weight: float, scalars: Tensor, pre: Tensor, act: Tensor, logits: Tensor,
                 row_lse: Tensor, row_ref: Tensor, grad_seq: Tensor, grad_scale: Tensor, need_dx: bool,
                 flat_grads: Tensor, simpo: bool, seg_rows: int = 0, seg_off: int = 0) -> Tensor:
    return head_bwd_impl(x, w1, b1, w2, b2, labels, seq_off, average, sft_weight, scalars, pre, act, logits, row_lse,
                         row_ref, grad_seq, grad_scale, need_dx, flat_grads, simpo, seg_rows, seg_off)


head_bwd = torch.library.custom_op("ospo_head::head_bwd", _head_bwd_op, mutates_args=("flat_grads",),
                                   device_types="cuda")
def _cfg_sample_op(h: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, cfg_weight: float, temperature: float,
                   uniforms: Tensor, greedy: bool, merge_mode: int, want_logits: bool = False) -> List[Tensor]:
    return cfg_sample_impl(h, w1, b1, w2, b2, cfg_weight, temperature, uniforms, greedy, merge_mode, want_logits)


cfg_sample = torch.library.custom_op("ospo_head::cfg_sample", _cfg_sample_op, mutates_args=(), device_types="cuda")
cfg_merge_sample = torch.library.custom_op("ospo_head::cfg_merge_sample", cfg_merge_sample_impl, mutates_args=(),
                                           device_types="cuda")


def _gen_img_embeds_op(ids: Tensor, gen_embed: Tensor, wa: Tensor, ba: Tensor, wb: Tensor, bb: Tensor,
                       repeat: int = 1) -> Tensor:
    return gen_img_embeds_impl(ids, gen_embed, wa, ba, wb, bb, repeat)


gen_img_embeds = torch.library.custom_op("ospo_head::gen_img_embeds", _gen_img_embeds_op, mutates_args=(),
                                         device_types="cuda")


@linear_gelu_linear.register_fake
def _(x, w1, b1, w2, b2):
    return x.new_empty(x.shape[0], w2.shape[0], dtype=torch.bfloat16)
