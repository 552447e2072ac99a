"""Oracle (test infrastructure only): Product2Vec arithmetic in numpy.

Restates /root/reference/src/models/product2vec.py.  Parameters are passed as a dict keyed
exactly like the reference ``state_dict`` (``ffn.0.weight`` ... ``attention.out_proj.bias``)
so the same weights drive the reference, this oracle and the CUDA path.

All functions compute in the dtype of their inputs (tests pass float64 for a tight truth,
float32 to mimic the reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

BN_EPS = 1e-5        # nn.BatchNorm1d default, product2vec.py:16
BN_MOMENTUM = 0.1    # nn.BatchNorm1d default
PAIRWISE_EPS = 1e-6  # F.pairwise_distance default eps, product2vec.py:137


# --------------------------------------------------------------------------- FFN
def linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    """nn.Linear: y = x W^T + b."""
    y = x @ w.T
    return y if b is None else y + b


def batchnorm1d(x, gamma, beta, running_mean, running_var, training: bool,
                eps: float = BN_EPS, momentum: float = BN_MOMENTUM):
    """nn.BatchNorm1d over rows (product2vec.py:16).

    Training: normalise with the biased batch variance, update running stats with the
    unbiased one (torch semantics).  Returns (y, new_running_mean, new_running_var).
    """
    if training:
        n = x.shape[0]
        mean = x.mean(axis=0)
        var = x.var(axis=0)  # biased
        y = (x - mean) / np.sqrt(var + eps) * gamma + beta
        unbiased = var * (n / max(n - 1, 1))
        new_rm = (1 - momentum) * running_mean + momentum * mean
        new_rv = (1 - momentum) * running_var + momentum * unbiased
        return y, new_rm, new_rv
    y = (x - running_mean) / np.sqrt(running_var + eps) * gamma + beta
    return y, running_mean, running_var


def ffn(params: Dict[str, np.ndarray], x: np.ndarray, training: bool = False):
    """The FFN of product2vec.py:14-21 on 2-D rows: Linear-BN-Tanh-Linear-Tanh-Linear.

    Returns (y, (new_running_mean, new_running_var)).
    """
    z = linear(x, params["ffn.0.weight"], params["ffn.0.bias"])
    z, rm, rv = batchnorm1d(z, params["ffn.1.weight"], params["ffn.1.bias"],
                            params["ffn.1.running_mean"], params["ffn.1.running_var"], training)
    z = np.tanh(z)
    z = np.tanh(linear(z, params["ffn.3.weight"], params["ffn.3.bias"]))
    return linear(z, params["ffn.5.weight"], params["ffn.5.bias"]), (rm, rv)


def get_initial_embedding(params, features: np.ndarray, training: bool = False):
    """product2vec.py:31-46: 1-D / 2-D / 3-D handling (3-D rows are flattened, so BatchNorm
    sees B*N rows), ValueError otherwise."""
    if features.ndim == 1:
        y, st = ffn(params, features[None, :], training)
        return y[0], st
    if features.ndim == 2:
        return ffn(params, features, training)
    if features.ndim == 3:
        b, n, d = features.shape
        y, st = ffn(params, features.reshape(-1, d), training)
        return y.reshape(b, n, -1), st
    raise ValueError(f"Unexpected input dimension: {features.ndim}")


# --------------------------------------------------------------------------- attention
def in_projection(params, h_query: np.ndarray, h_kv: np.ndarray):
    """Packed in-projection of nn.MultiheadAttention for query!=key, key is value
    (torch F._in_projection_packed): Q from rows [0:E] of in_proj_weight, K|V from [E:3E]."""
    w, b = params["attention.in_proj_weight"], params["attention.in_proj_bias"]
    e = w.shape[1]
    q = linear(h_query, w[:e], b[:e])
    kv = linear(h_kv, w[e:], b[e:])
    return q, kv  # kv[..., :E] = K, kv[..., E:] = V


def mha_dense(params, query: np.ndarray, key_value: np.ndarray, heads: int) -> np.ndarray:
    """nn.MultiheadAttention(query[B,E] as one target position, key_value[B,N,E]) in eval /
    dropout-0 mode, need_weights branch: softmax_j((q*sqrt(1/dh)) . k_j) V, out_proj.
    product2vec.py:48-68."""
    b, n, e = key_value.shape
    dh = e // heads
    q, kv = in_projection(params, query, key_value)
    k, v = kv[..., :e], kv[..., e:]
    qh = (q * math.sqrt(1.0 / dh)).reshape(b, heads, dh)
    kh = k.reshape(b, n, heads, dh)
    vh = v.reshape(b, n, heads, dh)
    s = np.einsum("bhd,bnhd->bhn", qh, kh)
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(axis=-1, keepdims=True)
    o = np.einsum("bhn,bnhd->bhd", p, vh).reshape(b, e)
    return linear(o, params["attention.out_proj.weight"], params["attention.out_proj.bias"])


def gat_csr_forward(q: np.ndarray, kv: np.ndarray, rowptr: np.ndarray, col: np.ndarray,
                    heads: int) -> Tuple[np.ndarray, np.ndarray]:
    """Attention core over a CSR: row i attends to kv[col[rowptr[i]:rowptr[i+1]]].

    q [n_dst, E] (already projected, NOT yet scaled), kv [n_src, 2E] (K | V).
    Returns (o [n_dst, E], lse [n_dst, heads]) with lse = log sum_j exp(s_ij) (natural log);
    rows without neighbours give o = 0, lse = 0.
    """
    n_dst, e = q.shape
    dh = e // heads
    scale = math.sqrt(1.0 / dh)
    o = np.zeros_like(q)
    lse = np.zeros((n_dst, heads), dtype=q.dtype)
    for i in range(n_dst):
        beg, end = int(rowptr[i]), int(rowptr[i + 1])
        if end == beg:
            continue
        nb = col[beg:end]
        k = kv[nb, :e].reshape(-1, heads, dh)
        v = kv[nb, e:].reshape(-1, heads, dh)
        s = np.einsum("hd,nhd->hn", (q[i] * scale).reshape(heads, dh), k)
        m = s.max(axis=-1, keepdims=True)
        p = np.exp(s - m)
        l = p.sum(axis=-1, keepdims=True)
        o[i] = np.einsum("hn,nhd->hd", p / l, v).reshape(e)
        lse[i] = (m + np.log(l))[:, 0]
    return o, lse


def gat_csr_backward(q, kv, rowptr, col, heads: int, d_o: np.ndarray):
    """Analytic gradient of gat_csr_forward w.r.t. q and kv given d_o (what autograd's
    BmmBackward/SoftmaxBackward/MulBackward produce for product2vec.py:60).

    dS_ij = a_ij (dO_i.V_j - dO_i.O_i);  dQ_i = scale * sum_j dS_ij K_j;
    dK_j += scale * dS_ij Q_i;  dV_j += a_ij dO_i.
    """
    n_dst, e = q.shape
    dh = e // heads
    scale = math.sqrt(1.0 / dh)
    dq = np.zeros_like(q)
    dkv = np.zeros_like(kv)
    for i in range(n_dst):
        beg, end = int(rowptr[i]), int(rowptr[i + 1])
        if end == beg:
            continue
        nb = col[beg:end]
        k = kv[nb, :e].reshape(-1, heads, dh)
        v = kv[nb, e:].reshape(-1, heads, dh)
        qi = q[i].reshape(heads, dh)
        s = np.einsum("hd,nhd->hn", qi * scale, k)
        s = s - s.max(axis=-1, keepdims=True)
        a = np.exp(s)
        a = a / a.sum(axis=-1, keepdims=True)              # [h, n]
        go = d_o[i].reshape(heads, dh)
        o = np.einsum("hn,nhd->hd", a, v)
        da = np.einsum("hd,nhd->hn", go, v)
        delta = (go * o).sum(axis=-1, keepdims=True)       # [h, 1]
        ds = a * (da - delta)
        dq[i] = (scale * np.einsum("hn,nhd->hd", ds, k)).reshape(e)
        dk = scale * np.einsum("hn,hd->nhd", ds, qi).reshape(-1, e)
        dv = np.einsum("hn,hd->nhd", a, go).reshape(-1, e)
        np.add.at(dkv, (nb, slice(0, e)), dk)
        np.add.at(dkv, (nb, slice(e, 2 * e)), dv)
    return dq, dkv


def apply_attention(params, query: np.ndarray, key_value: np.ndarray, heads: int) -> np.ndarray:
    """product2vec.py:48-68 including its 1-D / 2-D dimension juggling."""
    q, kv = query, key_value
    if q.ndim == 1:
        q = q[None, :]
    if kv.ndim == 2:
        kv = kv[None, :, :]
    out = mha_dense(params, q, kv, heads)
    if query.ndim == 1:
        return out[0]
    return out


def forward(params, features: np.ndarray, neighbors: Optional[np.ndarray], heads: int,
            training: bool = False):
    """Product2Vec.forward (product2vec.py:70-81).  Dropout on the attention weights is not
    modelled (parity runs use DROPOUT=0 or eval mode).

    Returns (embeddings, [bn_stats_after_features, bn_stats_after_neighbors?]).
    In training mode the running statistics are threaded through the two FFN calls in the
    order the reference makes them (features first, then neighbours).
    """
    p = dict(params)
    emb, (rm, rv) = get_initial_embedding(p, features, training)
    p["ffn.1.running_mean"], p["ffn.1.running_var"] = rm, rv
    if neighbors is not None and neighbors.shape[0] > 0:
        nb, (rm, rv) = get_initial_embedding(p, neighbors, training)
        p["ffn.1.running_mean"], p["ffn.1.running_var"] = rm, rv
        emb = apply_attention(p, emb, nb, heads)
    return emb, (p["ffn.1.running_mean"], p["ffn.1.running_var"])


def forward_graph(params, x: np.ndarray, rowptr: np.ndarray, col: np.ndarray, heads: int,
                  training: bool = False, double_ffn_query: bool = False):
    """Full-graph formulation: FFN once per node over the [N, E] node rows
    (BatchNorm batch = the node rows), K|V per node, attention over the CSR, out-proj;
    nodes without out-neighbours keep ffn(x) (product2vec.py:76, :98).

    double_ffn_query=True reproduces generate_all_embeddings (product2vec.py:90-108):
    the query is ffn(ffn(x_i)) while keys/values are ffn(x_j).
    """
    h, (rm, rv) = ffn(params, x, training)
    p = dict(params)
    p["ffn.1.running_mean"], p["ffn.1.running_var"] = rm, rv
    hq = h
    if double_ffn_query:
        hq, (rm, rv) = ffn(p, h, training)
        p["ffn.1.running_mean"], p["ffn.1.running_var"] = rm, rv
    q, kv = in_projection(p, hq, h)
    o, _ = gat_csr_forward(q, kv, rowptr, col, heads)
    out = linear(o, p["attention.out_proj.weight"], p["attention.out_proj.bias"])
    deg = np.diff(rowptr)
    out = np.where((deg > 0)[:, None], out, h)
    return out, (rm, rv)


# --------------------------------------------------------------------------- loss
def pairwise_distance(a: np.ndarray, b: np.ndarray, eps: float = PAIRWISE_EPS) -> np.ndarray:
    """F.pairwise_distance(p=2): || a - b + eps ||_2 over the last axis."""
    d = a - b + eps
    return np.sqrt((d * d).sum(axis=-1))


def triplet_hinge(anchor, positive, negative, margin: float):
    """product2vec.py:137-154 (sign as written in the reference):
    mean_i relu(margin - d(a_i,p_i) + mean_k d(a_i,n_ik)).  Returns (loss, per_sample)."""
    dpos = pairwise_distance(anchor, positive)
    if negative.ndim == 3:
        dneg = pairwise_distance(anchor[:, None, :], negative).mean(axis=1)
    else:
        dneg = pairwise_distance(anchor, negative)
    per = np.maximum(margin - dpos + dneg, 0.0)
    return per.mean(), per


def triplet_hinge_backward(anchor, positive, negative, margin: float, grad_loss: float = 1.0):
    """Gradient of triplet_hinge w.r.t. (anchor, positive, negative[B,K,D])."""
    b = anchor.shape[0]
    kneg = negative.shape[1]
    dp = anchor - positive + PAIRWISE_EPS
    npos = np.sqrt((dp * dp).sum(-1))
    dn = anchor[:, None, :] - negative + PAIRWISE_EPS
    nneg = np.sqrt((dn * dn).sum(-1))
    active = (margin - npos + nneg.mean(1)) > 0
    g = np.where(active, grad_loss / b, 0.0)[:, None]
    up = dp / npos[:, None]
    un = dn / nneg[:, :, None]
    ga = g * (-up + un.mean(1))
    gp = g * up
    gn = -(g[:, None, :] * un) / kneg
    return ga, gp, gn
