"""Oracle (test infrastructure only): P-Companion joint model arithmetic in numpy.

Restates /root/reference/src/models/type_transition.py, item_prediction.py and
p_companion.py.  Parameters are keyed like the reference ``state_dict``.  Dropout
(type_transition.py:17) is not modelled: parity runs use eval mode or DROPOUT=0.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .p2v import linear


def type_transition(params: Dict[str, np.ndarray], t: np.ndarray, prefix: str = "type_transition.") -> np.ndarray:
    """type_transition.py:15-20: decoder(relu(encoder(t)))."""
    h = np.maximum(linear(t, params[prefix + "encoder.weight"], params[prefix + "encoder.bias"]), 0.0)
    return linear(h, params[prefix + "decoder.weight"], params[prefix + "decoder.bias"])


def item_prediction(params, q_item: np.ndarray, comp_type_emb: np.ndarray,
                    prefix: str = "item_prediction.") -> np.ndarray:
    """item_prediction.py:31-38: item_projection(q)[:, None, :] * type_projection(T)."""
    pi = linear(q_item, params[prefix + "item_projection.weight"], params[prefix + "item_projection.bias"])
    pt = linear(comp_type_emb, params[prefix + "type_projection.weight"], params[prefix + "type_projection.bias"])
    return pi[:, None, :] * pt


def topk_stable(scores: np.ndarray, k: int):
    """Row-wise top-k, descending score, ties -> lowest index (the contract the CUDA path
    implements; torch.topk itself leaves tie order unspecified)."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, order, axis=1), order.astype(np.int64)


def pcompanion_forward(params, query_idx: np.ndarray, query_types: np.ndarray, k_types: int):
    """p_companion.py:45-77 with integer query indices (the pid -> idx dict lookup of
    :47-49 is host-side)."""
    q_emb = params["product_embeddings.weight"][query_idx]
    t_emb = params["query_type_embeddings.weight"][query_types]
    base = type_transition(params, t_emb)
    sims = base @ params["complementary_type_embeddings.weight"].T
    _, top = topk_stable(sims, k_types)
    comp = params["complementary_type_embeddings.weight"][top]
    proj = item_prediction(params, q_emb, comp)
    return {"projected_embeddings": proj, "complementary_types": top, "type_similarities": sims}


def type_hinge(sims: np.ndarray, pos: np.ndarray, neg: np.ndarray, margin: float) -> float:
    """p_companion.py:95-103."""
    r = np.arange(sims.shape[0])
    return np.maximum(margin - sims[r, pos] + sims[r, neg], 0.0).mean()


def item_hinge(proj: np.ndarray, pos_items: np.ndarray, neg_items: np.ndarray, margin: float) -> float:
    """p_companion.py:105-119: torch.norm (no eps) over the last dim, mean over [B, K]."""
    dp = np.sqrt(((proj - pos_items[:, None, :]) ** 2).sum(-1))
    dn = np.sqrt(((proj - neg_items[:, None, :]) ** 2).sum(-1))
    return np.maximum(margin - dp + dn, 0.0).mean()


def item_hinge_backward(proj, pos_items, neg_items, margin: float, grad_loss: float = 1.0):
    """d item_hinge / d proj."""
    a = proj - pos_items[:, None, :]
    b = proj - neg_items[:, None, :]
    dp = np.sqrt((a * a).sum(-1))
    dn = np.sqrt((b * b).sum(-1))
    active = (margin - dp + dn) > 0
    g = np.where(active, grad_loss / dp.size, 0.0)[..., None]
    return g * (-a / dp[..., None] + b / dn[..., None])


def compute_loss(outputs, positive_types, negative_types, positive_items, negative_items,
                 alpha: float, margin: float) -> float:
    """p_companion.py:79-93: alpha * item + (1 - alpha) * type."""
    tl = type_hinge(outputs["type_similarities"], positive_types, negative_types, margin)
    il = item_hinge(outputs["projected_embeddings"], positive_items, negative_items, margin)
    return alpha * il + (1 - alpha) * tl


def hit_at_k(predictions: np.ndarray, ground_truth: np.ndarray, k: int) -> float:
    """metrics.py:7-26 with the stable tie rule."""
    k = min(k, predictions.shape[1])
    _, top = topk_stable(predictions, k)
    return float((top == ground_truth[:, None]).any(axis=1).mean())
