"""Scalable synthetic BPG generator (device side, O(E)).

The reference generator (/root/reference/src/data/synthetic_data.py:78-153) enumerates all
C(P, 2) pairs with itertools.combinations and cannot go beyond a few thousand products
(SURVEY H9).  This restates its distribution per *sampled* pair: the same category-biased
Bernoulli chain (co-view 0.3, x1.5 inside a category; purchase-after-view 0.2 given co-view;
co-purchase 0.15, x0.5 inside a category; synthetic_data.py:101-128), the same src < dst edge
convention (:94), and the same features (randn + 1.0 on the category's 20-wide slice, :50-52).
It only synthesises *inputs* for benchmarks and tests; the graph itself is then built by the
CUDA sort / unique / set kernels (BehaviorProductGraph.from_arrays).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .bpg import BehaviorProductGraph

NUM_CATEGORIES = 5  # synthetic_data.py:37


def edge_chain(category: torch.Tensor, src: torch.Tensor, dst: torch.Tensor, u_cv: torch.Tensor, u_pav: torch.Tensor,
               u_cp: torch.Tensor, co_view_prob: float = 0.3, pav_given_cv_prob: float = 0.2,
               cp_given_pav_prob: float = 0.15) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """synthetic_data.py:101-128 on pre-sampled candidate pairs and pre-drawn uniforms."""
    same = category[src.long()] == category[dst.long()]
    cv_prob = torch.where(same, co_view_prob * 1.5, co_view_prob)
    cp_prob = torch.where(same, cp_given_pav_prob * 0.5, cp_given_pav_prob)
    is_cv = u_cv < cv_prob
    is_pav = is_cv & (u_pav < pav_given_cv_prob)
    cp_in = is_pav & (u_cp < cp_prob)
    comp = (~is_cv) & (u_cp < cp_prob)
    is_cp = cp_in | comp
    pick = lambda m: (src[m].contiguous(), dst[m].contiguous())
    return {"co_view": pick(is_cv), "purchase_after_view": pick(is_pav), "co_purchase": pick(is_cp)}


def synthetic_catalog(num_products: int, num_types: int = 20, dim: int = 128, seed: int = 1234,
                      device: Optional[torch.device] = None):
    """(features [P, dim] fp32, type_id int32 [P], category int32 [P])."""
    device = torch.device(device or "cuda")
    g = torch.Generator(device=device).manual_seed(seed)
    type_id = torch.randint(0, num_types, (num_products,), generator=g, device=device, dtype=torch.int32)
    per_cat = max(num_types // NUM_CATEGORIES, 1)
    category = torch.clamp(type_id // per_cat, max=NUM_CATEGORIES - 1).to(torch.int32)
    feats = torch.randn(num_products, dim, generator=g, device=device, dtype=torch.float32)
    width = min(20, dim // NUM_CATEGORIES)
    cols = torch.arange(dim, device=device).unsqueeze(0)
    lo = (category.long() * width).unsqueeze(1)
    feats += ((cols >= lo) & (cols < lo + width)).float()
    return feats, type_id, category


def synthetic_bpg(num_products: int, target_co_view_edges: int, num_types: int = 20, dim: int = 128,
                  seed: int = 1234, device: Optional[torch.device] = None) -> BehaviorProductGraph:
    """Synthetic BPG with about `target_co_view_edges` co-view edges (configs C2 / C5)."""
    device = torch.device(device or "cuda")
    feats, type_id, category = synthetic_catalog(num_products, num_types, dim, seed, device)
    g = torch.Generator(device=device).manual_seed(seed + 1)
    n_pairs = int(target_co_view_edges / 0.33)
    a = torch.randint(0, num_products, (n_pairs,), generator=g, device=device, dtype=torch.int32)
    b = torch.randint(0, num_products, (n_pairs,), generator=g, device=device, dtype=torch.int32)
    keep = a != b
    src, dst = torch.minimum(a, b)[keep], torch.maximum(a, b)[keep]          # src < dst, :94
    u = torch.rand(3, src.numel(), generator=g, device=device)
    edges = edge_chain(category, src, dst, u[0], u[1], u[2])
    del a, b, keep, u
    return BehaviorProductGraph.from_arrays(num_products, edges, feats, type_id, device)
