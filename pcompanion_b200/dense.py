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
    return k in (128, 256) and n in (128, 256)


class _LinearTC(torch.autograd.Function):
    """y = act(x W^T + b) with act in {identity, tanh}; all three GEMMs on tcgen05."""

    @staticmethod
    def forward(ctx, x, weight, bias, tanh: bool):
        x = x if x.stride(-1) == 1 else x.contiguous()
        y = ops.linear_tc(x, weight.contiguous(), bias, ops.EPI_BIAS_TANH if tanh else ops.EPI_BIAS)
        ctx.save_for_backward(x, weight, y if tanh else None)
        ctx.tanh = tanh
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.tanh:
            dy = dy * (1.0 - y * y)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear_tc(dy, weight.t().contiguous(), None)          # dX = dY . W
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = ops.wgrad_tc(dy, x, want_bias=ctx.has_bias)
        return dx, dw, db, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], tanh: bool = False) -> torch.Tensor:
    """nn.Linear (optionally followed by tanh) on [rows, k] -> [rows, n]."""
    if not x.is_cuda:
        raise RuntimeError(f"pcompanion_b200.dense.linear: input on {x.device}; CUDA only (no CPU fallback)")
    n, k = weight.shape
    if x.dim() == 2 and _tc_shape_ok(k, n) and x.dtype == torch.float32:
        return _LinearTC.apply(x, weight, bias, tanh)
    y = F.linear(x, weight, bias)
    return torch.tanh(y) if tanh else y


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
    l0, bn, _, l3, _, l5 = ffn
    z = linear(rows, l0.weight, l0.bias)
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    z = F.batch_norm(z, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)
    z = torch.tanh(z)
    z = linear(z, l3.weight, l3.bias, tanh=True)
    return linear(z, l5.weight, l5.bias)
