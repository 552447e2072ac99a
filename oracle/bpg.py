"""Oracle (test infrastructure only): behaviour-product-graph integer logic.

Restates /root/reference/src/data/bpg.py (edge sets, neighbour lookup, the two set helpers)
and the inline set algebra of /root/reference/src/data/synthetic_data.py:89-90,110-128 on
integer node indices.  Results are canonical (sorted ascending by (src, dst)) so they can be
compared bit-exactly with the device CSR.
"""
from __future__ import annotations

from typing import Dict, Set, Tuple

import numpy as np

EDGE_TYPES = ("co_purchase", "co_view", "purchase_after_view")  # bpg.py:9-13


def pack_keys(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """(src, dst) -> uint64 key src<<32 | dst; sorting keys sorts by (src, dst)."""
    return (src.astype(np.uint64) << np.uint64(32)) | dst.astype(np.uint64)


def unpack_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    return (keys >> np.uint64(32)).astype(np.int32), (keys & np.uint64(0xFFFFFFFF)).astype(np.int32)


def unique_sorted_keys(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """Set semantics of ``edges[edge_type].add((s, t))`` (bpg.py:19-22): duplicates collapse."""
    return np.unique(pack_keys(np.asarray(src), np.asarray(dst)))


def csr_from_keys(keys: np.ndarray, num_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """Sorted-unique keys -> (rowptr int64[N+1], col int32[E]); row i lists, ascending, the
    targets t of edges (i, t) - i.e. ``get_neighbors(i, edge_type)`` (bpg.py:24-31)."""
    src, dst = unpack_keys(keys)
    counts = np.bincount(src, minlength=num_nodes).astype(np.int64)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, dst.astype(np.int32)


def csc_from_csr(rowptr: np.ndarray, col: np.ndarray, num_src: int):
    """Transpose: for every column node j the ascending list of rows i with an edge (i, j).
    Returns (colptr int64[num_src+1], row int32[E], perm int64[E]) with perm[e_csc] = e_csr."""
    n = len(rowptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rowptr))
    order = np.lexsort((rows, col))  # by col, then row: stable & canonical
    counts = np.bincount(col, minlength=num_src).astype(np.int64)
    colptr = np.zeros(num_src + 1, dtype=np.int64)
    np.cumsum(counts, out=colptr[1:])
    return colptr, rows[order].astype(np.int32), order.astype(np.int64)


def neighbors(rowptr: np.ndarray, col: np.ndarray, i: int) -> np.ndarray:
    return col[rowptr[i]:rowptr[i + 1]]


def set_intersection(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.intersect1d(a, b, assume_unique=True)


def set_difference(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.setdiff1d(a, b, assume_unique=True)


def set_union(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.union1d(a, b)


def similarity_pairs(cv: np.ndarray, pav: np.ndarray, cp: np.ndarray) -> np.ndarray:
    """(Bcv n Bpv) - Bcp, synthetic_data.py:89,118-119."""
    return set_difference(set_intersection(cv, pav), cp)


def complementary_pairs(cv: np.ndarray, pav: np.ndarray, cp: np.ndarray) -> np.ndarray:
    """Bcp - (Bpv u Bcv), synthetic_data.py:90,125-128."""
    return set_difference(cp, set_union(pav, cv))


def exclusive_co_purchase_pairs(cv: np.ndarray, cp: np.ndarray) -> np.ndarray:
    """bpg.py:51-56: cp - cv."""
    return set_difference(cp, cv)


def co_view_intersection_pairs(cv: np.ndarray, pav: np.ndarray) -> np.ndarray:
    """bpg.py:58-63: cv n pav."""
    return set_intersection(cv, pav)


# ---- pure-Python restatement on tuples (small cases; mirrors the reference's data types)
class SetBPG:
    """dict/set BPG exactly as bpg.py:7-38 stores it, over arbitrary hashable ids."""

    def __init__(self) -> None:
        self.nodes: Dict = {}
        self.edges: Dict[str, Set[tuple]] = {t: set() for t in EDGE_TYPES}

    def add_node(self, pid, features) -> None:
        self.nodes[pid] = features

    def add_edge(self, s, t, edge_type: str) -> None:
        if edge_type in self.edges:          # unknown types silently dropped, bpg.py:21
            self.edges[edge_type].add((s, t))

    def get_neighbors(self, pid, edge_type=None) -> set:
        out = set()
        if edge_type and edge_type in self.edges:
            out.update(t for s, t in self.edges[edge_type] if s == pid)
        else:
            for es in self.edges.values():
                out.update(t for s, t in es if s == pid)
        return out

    def get_all_types(self) -> set:
        return {n["type"] for n in self.nodes.values()}

    def get_products_by_type(self, product_type) -> list:
        return [pid for pid, d in self.nodes.items() if d["type"] == product_type]


def edge_chain(num_nodes: int, category: np.ndarray, pair_src: np.ndarray, pair_dst: np.ndarray,
               u_cv: np.ndarray, u_pav: np.ndarray, u_cp: np.ndarray,
               co_view_prob: float = 0.3, pav_given_cv_prob: float = 0.2,
               cp_given_pav_prob: float = 0.15) -> Dict[str, np.ndarray]:
    """The Bernoulli chain of synthetic_data.py:101-128 applied to pre-sampled candidate
    pairs with pre-drawn uniforms (O(E); the reference's O(P^2) itertools.combinations pair
    enumeration cannot scale).  Returns packed key arrays (not yet deduplicated)
    for the three edge types plus the generator's own similarity / complementary sets.
    """
    same = category[pair_src] == category[pair_dst]
    cv_prob = np.where(same, co_view_prob * 1.5, co_view_prob)
    cp_prob = np.where(same, cp_given_pav_prob * 0.5, cp_given_pav_prob)
    is_cv = u_cv < cv_prob
    is_pav = is_cv & (u_pav < pav_given_cv_prob)
    # synthetic_data.py:118-122: inside the pav branch, u_cp >= cp_prob -> similarity, else cp
    sim = is_pav & (u_cp >= cp_prob)
    cp_in = is_pav & (u_cp < cp_prob)
    # synthetic_data.py:123-128: not co-viewed, u_cp < cp_prob -> cp and complementary
    comp = (~is_cv) & (u_cp < cp_prob)
    keys = pack_keys(pair_src, pair_dst)
    return {
        "co_view": keys[is_cv],
        "purchase_after_view": keys[is_pav],
        "co_purchase": keys[cp_in | comp],
        "similarity": keys[sim],
        "complementary": keys[comp],
    }
