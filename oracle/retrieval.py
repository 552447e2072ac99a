"""Oracle (test infrastructure only): complementary retrieval (masked top-K).

Restates the retrieval semantics of /root/reference/inference.py:93-113 (for every predicted
complementary type: restrict the catalog to products of that type, score = projected . item,
keep the top-k) and of /root/reference/src/utils/metrics.py:89-100 (unmasked in-batch scoring),
generalised to one masked top-K over the whole catalog as BASELINE.json's north_star asks.

Score definition (shared with the CUDA path so indices can be compared bit-exactly):
    score[r, p] = sum_d double(q[r, d]) * double(c[p, d])
accumulated in float64 sequentially over d = 0 .. D-1 (the order a thread of the CUDA kernel uses).  Each product of two float32 values is exact in float64, so a fused
multiply-add and a separate multiply/add round identically; only the summation order matters
and it is fixed.  Ranking: descending score, ties -> lowest catalog index (a stable sort);
rows with fewer than k eligible products are padded with index -1 / score -inf.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def scores_fp64(q: np.ndarray, catalog: np.ndarray) -> np.ndarray:
    """[R, D] x [P, D] -> [R, P] float64, accumulated sequentially over d = 0 .. D-1 (the order one
    thread of the CUDA kernel uses for one (row, product) pair)."""
    q64 = q.astype(np.float64)
    c64 = catalog.astype(np.float64)
    acc = np.zeros((q.shape[0], catalog.shape[0]), dtype=np.float64)
    for d in range(q.shape[1]):
        acc += q64[:, d:d + 1] * c64[None, :, d]
    return acc


def masked_topk(q: np.ndarray, catalog: np.ndarray, k: int,
                row_type: Optional[np.ndarray] = None, type_id: Optional[np.ndarray] = None,
                index_base: int = 0, chunk: int = 1 << 16) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of every row against the catalog, restricted to type_id[p] == row_type[r] when
    both are given (row_type < 0 means "no restriction" for that row).

    Returns (scores float64 [R, k], indices int64 [R, k]); indices are global (index_base is
    added, for sharded catalogs).
    """
    r = q.shape[0]
    best_s = np.full((r, k), -np.inf)
    best_i = np.full((r, k), -1, dtype=np.int64)
    for beg in range(0, catalog.shape[0], chunk):
        end = min(beg + chunk, catalog.shape[0])
        s = scores_fp64(q, catalog[beg:end])
        if row_type is not None and type_id is not None:
            ok = (type_id[None, beg:end] == row_type[:, None]) | (row_type[:, None] < 0)
            s = np.where(ok, s, -np.inf)
        idx = np.broadcast_to(np.arange(beg, end, dtype=np.int64) + index_base, s.shape)
        cat_s = np.concatenate([best_s, s], axis=1)
        cat_i = np.concatenate([best_i, idx], axis=1)
        cat_i = np.where(np.isneginf(cat_s), -1, cat_i)
        best_s, best_i = merge_topk(cat_s, cat_i, k)
    return best_s, best_i


def merge_topk(scores: np.ndarray, indices: np.ndarray, k: int):
    """Merge candidate lists (e.g. per-shard top-K): order by (score desc, index
    asc) with padding entries (index -1) last; keep k."""
    pad = indices < 0
    key_idx = np.where(pad, np.iinfo(np.int64).max, indices)
    order = np.lexsort((key_idx, -scores), axis=1)[:, :k]
    return np.take_along_axis(scores, order, axis=1), np.take_along_axis(indices, order, axis=1)
