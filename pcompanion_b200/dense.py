"""Dense per-row projections of the hot path (FFN chain, Q / K|V / out projections, the small
type-transition and item-prediction layers).

These are the only GEMM-shaped operations on the path.  They are
plain library GEMMs: torch dispatches them to cuBLAS in full fp32 (TF32 is left disabled so the
1e-5 parity gate of BASELINE.json holds), with BatchNorm / tanh as ATen elementwise kernels.
Everything irregular - gathers, segmented softmax, scatter, hinge reductions, sort / set logic,
top-K - is hand-written CUDA behind the C ABI.  CUDA tensors only: there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError(f"pcompanion_b200.dense.linear: input on {x.device}; CUDA only (no CPU fallback)")
    return F.linear(x, weight, bias)


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
    z = torch.tanh(linear(z, l3.weight, l3.bias))
    return linear(z, l5.weight, l5.bias)
