"""TEST INFRASTRUCTURE -- chunked fp32 restatement of the SimPO head step in plain PyTorch on the GPU.

The CPU oracle (oracle/head_oracle.py) needs ~1 min for BASELINE.json's full configs[1] (37 TFLOP in fp32), so the
full-size parity tests use this restatement instead: the same reference lines (janus/models/modeling_vlm.py:47-51,
ospo/wrapper/train.py:317-342, 375-396, 419) evaluated row-chunk by row-chunk in fp32 with aten ops
(``allow_tf32`` off), the pair-level stage through the oracle's own ``simpo_loss`` + autograd, and the backward of
the head as the analytic formulas of SURVEY §8 a-6.  ``tests/test_gpu_head.py`` checks it against the CPU oracle at
a size both can run, then uses it at 64 and 128 pairs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import head_oracle as O


@torch.no_grad()
def simpo_step_chunked_fp32(W1, b1, W2, b2, hidden, labels, T, L, *, beta, gamma_beta_ratio, label_smoothing=0.0,
                            loss_type="sigmoid", chunk_rows=4608, need_grads=True):
    """hidden [2B, L+T, H] (any float dtype, CUDA), labels [2B, L+T]; weights fp32 CUDA.
    Rows L-1 .. L+T-2 of every sequence predict the T image ids (train.py:385-387)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        S = hidden.shape[0]
        B = S // 2
        H = hidden.shape[2]
        dev = hidden.device
        x_rows = hidden[:, L - 1:L - 1 + T, :]                       # [S, T, H] view
        tgt = labels[:, L:L + T].reshape(-1)                          # [S*T]
        N = S * T
        per_seq = max(1, chunk_rows // T)
        row_logps = torch.empty(N, dtype=torch.float32, device=dev)
        row_lse = torch.empty(N, dtype=torch.float32, device=dev)
        for s0 in range(0, S, per_seq):
            s1 = min(S, s0 + per_seq)
            x = x_rows[s0:s1].reshape(-1, H).float()
            logits = F.linear(F.gelu(F.linear(x, W1, b1)), W2, b2)
            lse = torch.logsumexp(logits, -1)
            t = tgt[s0 * T:s1 * T]
            row_lse[s0 * T:s1 * T] = lse
            row_logps[s0 * T:s1 * T] = logits.gather(1, t[:, None]).squeeze(1) - lse
        with torch.enable_grad():
            seq = row_logps.view(S, T).mean(-1).detach().requires_grad_(True)       # train.py:394
            losses, crew, rrew = O.simpo_loss(seq[:B], seq[B:], beta, gamma_beta_ratio, label_smoothing, loss_type)
            loss = losses.mean()                                                    # train.py:419
            (gseq,) = torch.autograd.grad(loss, seq)
        out = {"loss": loss.detach(), "chosen_logps": seq[:B].detach(), "rejected_logps": seq[B:].detach(),
               "per_token_logps": row_logps}
        if not need_grads:
            return out
        V, E = W2.shape
        dW2 = torch.zeros(V, E, dtype=torch.float32, device=dev)
        dW1 = torch.zeros(E, H, dtype=torch.float32, device=dev)
        db2 = torch.zeros(V, dtype=torch.float32, device=dev)
        db1 = torch.zeros(E, dtype=torch.float32, device=dev)
        dx = torch.zeros(hidden.shape, dtype=torch.float32, device=dev)
        for s0 in range(0, S, per_seq):
            s1 = min(S, s0 + per_seq)
            x = x_rows[s0:s1].reshape(-1, H).float()
            pre = F.linear(x, W1, b1)
            act = F.gelu(pre)
            logits = F.linear(act, W2, b2)
            p = torch.softmax(logits, -1)
            t = tgt[s0 * T:s1 * T]
            c = (gseq[s0:s1] / T).repeat_interleave(T)
            dlog = -p * c[:, None]
            dlog[torch.arange(dlog.shape[0], device=dev), t] += c
            dW2 += dlog.t() @ act
            db2 += dlog.sum(0)
            dact = dlog @ W2
            cdf = 0.5 * (1.0 + torch.erf(pre * 0.7071067811865476))
            pdf = torch.exp(-0.5 * pre * pre) * 0.3989422804014327
            dpre = dact * (cdf + pre * pdf)
            dW1 += dpre.t() @ x
            db1 += dpre.sum(0)
            dx[s0:s1, L - 1:L - 1 + T, :] = (dpre @ W1).view(s1 - s0, T, H)
        out.update(dx=dx, dW2=dW2, dW1=dW1, db2=db2, db1=db1)
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
