"""Dense per-row projections of the hot path (FFN chain, Q / K|V / out projections).

The Product2Vec projections (128/256/384-wide) run on the tensor cores
through pcompanion_b200/csrc/gemm.cu: tcgen05.mma kind::tf32 with a 3-way hi/lo operand split
(fp32-faithful, so the 1e-5 parity gate holds), TMA-fed, accumulators in TMEM; forward, input
gradient (same kernel on the transposed weight) and weight / bias gradient (MN-major operands,
deterministic split reduction).  Layers whose shape the kernels are not instantiated for (the
64 -> 32 -> 64 type-transition MLP and the 64 -> 128 type projection of P-Companion, a few
hundred rows per step) are plain library GEMMs (cuBLAS via torch, full fp32).  BatchNorm
statistics / normalisation and tanh are elementwise passes.  CUDA tensors only.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import ops


def _tc_shape_ok(k: int, n: int) -> bool:
    """nn.Linear(k, n) whose forward, input gradient and weight gradient all fit the tcgen05 kernels: the forward / dgrad
    kernel needs both widths to be multiples of 32 (<= 768), the wgrad kernel n % 128 == 0 and k <= 256."""
    return k % 32 == 0 and 32 <= k <= 256 and n % 128 == 0 and n <= 768


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b; forward, input gradient and weight / bias gradient on tcgen05."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x if x.stride(-1) == 1 else x.contiguous()
        y = ops.linear_tc(x, weight.contiguous(), bias, ops.EPI_BIAS)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear_tc(dy, weight.t().contiguous(), None)          # dX = dY . W
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = ops.wgrad_tc(dy, x, want_bias=ctx.has_bias)
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.Linear on [rows, k] -> [rows, n]."""
    if not x.is_cuda:
        raise RuntimeError(f"pcompanion_b200.dense.linear: input on {x.device}; CUDA only (no CPU fallback)")
    n, k = weight.shape
    if x.dim() == 2 and _tc_shape_ok(k, n) and x.dtype == torch.float32 and x.shape[0] > 0:
        return _LinearTC.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class _TypeScoresTC(torch.autograd.Function):
    """S = base . W^T for a wide type table (P-Companion's [B, L] x [L, T] scoring, p_companion.py:60-62, T = 34,800 by
    default) on the tcgen05 kernel: column chunks of 768 straight into the strided [B, T] output; the last T % 32
    columns are a small library GEMM.  The type loss does not differentiate through S (ops.type_hinge takes the
    factors); if a caller does, the dense gradients are library GEMMs."""
    CHUNK = 768

    @staticmethod
    def forward(ctx, base, weight):
        base = base.contiguous()
        weight = weight.contiguous()
        b, t = base.shape[0], weight.shape[0]
        out = torch.empty(b, t, dtype=torch.float32, device=base.device)
        main = t - t % 32
        for c0 in range(0, main, _TypeScoresTC.CHUNK):
            c1 = min(c0 + _TypeScoresTC.CHUNK, main)
            ops.linear_tc(base, weight[c0:c1], None, out0=out[:, c0:c1])
        if main < t:
            out[:, main:] = base @ weight[main:].t()
        ctx.save_for_backward(base, weight)
        return out

    @staticmethod
    def backward(ctx, d_out):
        base, weight = ctx.saved_tensors
        d_base = d_out @ weight if ctx.needs_input_grad[0] else None
        d_w = d_out.t() @ base if ctx.needs_input_grad[1] else None
        return d_base, d_w


class _TypeScoresTopK(torch.autograd.Function):
    """(S = base . W^T, top-k columns of every row): one tcgen05 GEMM with the top-k in its epilogue
    (pc_type_scores_topk).  The type loss does not differentiate through S (ops.type_hinge takes the factors); if a
    caller does, the dense gradients are library GEMMs."""

    @staticmethod
    def forward(ctx, base, weight, k: int):
        sims, _, top = ops.type_scores_topk(base, weight, k, materialize=True)
        ctx.save_for_backward(base, weight)
        ctx.mark_non_differentiable(top)
        return sims, top

    @staticmethod
    def backward(ctx, d_sims, _d_top):
        base, weight = ctx.saved_tensors
        d_base = d_sims @ weight if ctx.needs_input_grad[0] else None
        d_w = d_sims.t() @ base if ctx.needs_input_grad[1] else None
        return d_base, d_w, None


def type_scores_topk(base: torch.Tensor, weight: torch.Tensor, k: int):
    """[B, L] x [T, L]^T -> ([B, T] scores, [B, k] int64 best columns, ties -> lowest column)   (p_companion.py:60-64)."""
    if not base.is_cuda:
        raise RuntimeError(f"pcompanion_b200.dense.type_scores_topk: input on {base.device}; CUDA only (no CPU fallback)")
    if ops.type_scores_topk_supported(base, weight, k):
        return _TypeScoresTopK.apply(base, weight, k)
    sims = type_scores(base, weight)                # shapes outside the kernel's range: library GEMM + row top-k kernel
    _, top = ops.topk_rows(sims.detach(), k)
    return sims, top


def type_scores(base: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """[B, L] x [T, L]^T -> [B, T].  Large batches run on the tensor-core kernel; below 16,384 rows the 2 * T / 768 launches
    cost more than the GEMM and a single library call is used."""
    if not base.is_cuda:
        raise RuntimeError(f"pcompanion_b200.dense.type_scores: input on {base.device}; CUDA only (no CPU fallback)")
    k = base.shape[1]
    if base.dtype == torch.float32 and base.shape[0] >= 16384 and k % 32 == 0 and k <= 256 and weight.shape[0] >= 1024 \
            and (weight.shape[0] * 4) % 16 == 0:
        return _TypeScoresTC.apply(base, weight)
    return F.linear(base, weight, None)


def in_projection(h_query: torch.Tensor, h_kv: torch.Tensor, w: torch.Tensor, b: torch.Tensor):
    """Packed in-projection of nn.MultiheadAttention for query != key, key is value
    (torch F._in_projection_packed): Q from rows [0:E] of in_proj_weight, K|V from rows [E:3E]."""
    e = w.shape[1]
    q = linear(h_query, w[:e], b[:e])
    kv = linear(h_kv, w[e:], b[e:])
    return q, kv


def ffn_forward(ffn: torch.nn.Sequential, rows: torch.Tensor, training: bool) -> torch.Tensor:
    """Linear -> BatchNorm1d -> Tanh -> Linear -> Tanh -> Linear on [rows, D] (product2vec.py:14-21).
    BatchNorm uses the batch statistics of exactly these rows in training mode and updates the
    running statistics with torch's momentum / unbiased-variance rule."""
    if not rows.is_cuda:
        raise RuntimeError(f"pcompanion_b200.dense.ffn_forward: input on {rows.device}; CUDA only (no CPU fallback)")
    from .fused import ffn_rows, ffn_supported
    if ffn_supported(ffn, rows):
        return ffn_rows(ffn, rows, training)      # one autograd node, every row-sized op in a C-ABI kernel
    # shapes the tcgen05 kernels are not instantiated for (non-default HIDDEN_SIZE / PRODUCT_EMB_DIM): library path
    l0, bn, _, l3, _, l5 = ffn
    z = linear(rows, l0.weight, l0.bias)
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = bn.momentum
    if momentum is None:     # nn.BatchNorm1d: cumulative moving average
        momentum = 1.0 / float(bn.num_batches_tracked) if (training and bn.num_batches_tracked is not None) else 0.0
    z = F.batch_norm(z, bn.running_mean, bn.running_var, bn.weight, bn.bias, training or not bn.track_running_stats, momentum, bn.eps)
    z = torch.tanh(z)
    z = torch.tanh(linear(z, l3.weight, l3.bias))
    return linear(z, l5.weight, l5.bias)
