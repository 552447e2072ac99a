"""Behaviour product graph with a device-resident CSR.

Drop-in for /root/reference/src/data/bpg.py: same class name, attributes (``nodes``, ``edges``)
and methods (``add_node``, ``add_edge``, ``get_neighbors``, ``get_all_types``,
``get_products_by_type``, ``get_exclusive_co_purchase_pairs``, ``get_co_view_intersection_pairs``)
with the same return types, so ``data_loader.py`` / ``train.py`` of the reference keep working.
Behind that surface the graph is held as int32 edge arrays and, once ``finalize()`` has run,
as sorted-unique 64-bit edge keys + CSR per edge type on the GPU (pcompanion_b200/csrc/bpg.cu):

* ``add_edge``'s set-insert deduplication (bpg.py:21)      -> radix sort + unique
* the O(E) list-comprehension scan of ``get_neighbors`` (:24-38) -> one CSR row slice
* the set algebra of synthetic_data.py:89-90,118-128 / bpg.py:51-63 -> sorted-set kernels

Large graphs skip the Python dict/set surface entirely via ``from_arrays``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Set, Tuple

import torch

from . import ops

EDGE_TYPES = ("co_purchase", "co_view", "purchase_after_view")  # bpg.py:9-13


class BehaviorProductGraph:
    """Behavior Product Graph (reference API) backed by device CSRs."""

    def __init__(self, device: Optional[torch.device] = None):
        self.nodes: Dict[str, Dict[str, Any]] = {}
        self.edges: Dict[str, Set[tuple]] = {t: set() for t in EDGE_TYPES}
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        # integer side (filled by finalize() or from_arrays())
        self._ids: Optional[List[str]] = None
        self._index: Optional[Dict[str, int]] = None
        self._num_nodes = 0
        self._pending: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._keys: Dict[str, torch.Tensor] = {}
        self._csr: Dict[str, ops.CSRGraph] = {}
        self._type_names: Optional[List[Any]] = None
        self.type_id: Optional[torch.Tensor] = None       # int32 [P] on device
        self.features: Optional[torch.Tensor] = None      # fp32 [P, D] on device
        self._type_members: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._dirty = True

    # ------------------------------------------------------------------ reference surface
    def add_node(self, product_id: str, features: Dict[str, Any]) -> None:
        self.nodes[product_id] = features
        self._dirty = True

    def add_edge(self, source_id: str, target_id: str, edge_type: str) -> None:
        if edge_type in self.edges:  # unknown edge types are silently dropped, bpg.py:21
            self.edges[edge_type].add((source_id, target_id))
            self._dirty = True

    def get_neighbors(self, product_id: str, edge_type: Optional[str] = None) -> Set[str]:
        """Out-neighbours of `product_id` in one edge type, or in the union of all (bpg.py:24-38)."""
        self.finalize()
        i = self._index.get(product_id) if self._index is not None else None
        if i is None:
            return set()
        if edge_type and edge_type in EDGE_TYPES:
            idx = self.neighbor_indices(i, edge_type)
        else:
            idx = torch.unique(torch.cat([self.neighbor_indices(i, t) for t in EDGE_TYPES]))
        return {self._ids[j] for j in idx.tolist()}

    def get_all_types(self) -> Set[str]:
        return {node["type"] for node in self.nodes.values()}

    def get_products_by_type(self, product_type: str) -> List[str]:
        """Products of one type in node insertion order (bpg.py:44-49)."""
        self.finalize()
        if self._type_names is None or product_type not in self._type_names:
            return []
        t = self._type_names.index(product_type)
        members, offsets = self.type_members()
        lo, hi = int(offsets[t].item()), int(offsets[t + 1].item())
        return [self._ids[j] for j in members[lo:hi].tolist()]

    def get_exclusive_co_purchase_pairs(self) -> List[tuple]:
        """co_purchase - co_view with label +1 (bpg.py:51-56)."""
        self.finalize()
        keys = ops.set_difference(self.keys("co_purchase"), self.keys("co_view"))
        return [(s, t, 1) for s, t in self._pairs_to_ids(keys)]

    def get_co_view_intersection_pairs(self) -> List[tuple]:
        """co_view n purchase_after_view with label -1 (bpg.py:58-63)."""
        self.finalize()
        keys = ops.set_intersection(self.keys("co_view"), self.keys("purchase_after_view"))
        return [(s, t, -1) for s, t in self._pairs_to_ids(keys)]

    # ------------------------------------------------------------------ integer / device side
    @classmethod
    def from_arrays(cls, num_nodes: int, edges: Dict[str, Tuple[torch.Tensor, torch.Tensor]],
                    features: Optional[torch.Tensor] = None, type_id: Optional[torch.Tensor] = None,
                    device: Optional[torch.device] = None) -> "BehaviorProductGraph":
        """Build from int32 (src, dst) device arrays per edge type (duplicates allowed) - the path
        for graphs too large for Python sets (configs C2-C5)."""
        g = cls(device)
        g._num_nodes = int(num_nodes)
        for t, (s, d) in edges.items():
            if t not in EDGE_TYPES:
                continue  # same silent drop as add_edge
            g._pending[t] = (s.to(g.device, torch.int32).contiguous(), d.to(g.device, torch.int32).contiguous())
        g.features = None if features is None else features.to(g.device, torch.float32).contiguous()
        g.type_id = None if type_id is None else type_id.to(g.device, torch.int32).contiguous()
        g._build_device()
        g._dirty = False
        return g

    @property
    def num_nodes(self) -> int:
        return self._num_nodes

    def finalize(self) -> "BehaviorProductGraph":
        """Index the dict/set surface and (re)build the device CSRs if anything changed."""
        if not self._dirty:
            return self
        if self.nodes:
            self._ids = list(self.nodes.keys())
            self._index = {p: i for i, p in enumerate(self._ids)}
            self._num_nodes = len(self._ids)
            for t in EDGE_TYPES:
                es = self.edges[t]
                idx = self._index
                pairs = [(idx[s], idx[d]) for s, d in es if s in idx and d in idx]
                arr = torch.tensor(pairs, dtype=torch.int32).reshape(-1, 2)
                self._pending[t] = (arr[:, 0].contiguous().to(self.device), arr[:, 1].contiguous().to(self.device))
            feats = [n.get("features") for n in self.nodes.values()]
            if feats and all(isinstance(f, torch.Tensor) for f in feats):
                self.features = torch.stack([f.detach().float().cpu() for f in feats]).to(self.device)
            types = [n.get("type") for n in self.nodes.values()]
            if types and all(t is not None for t in types):
                self._type_names = sorted(set(types), key=str)
                lut = {t: i for i, t in enumerate(self._type_names)}
                self.type_id = torch.tensor([lut[t] for t in types], dtype=torch.int32, device=self.device)
        self._build_device()
        self._dirty = False
        return self

    def _build_device(self) -> None:
        self._keys.clear()
        self._csr.clear()
        self._type_members = None
        n = self._num_nodes
        for t in EDGE_TYPES:
            s, d = self._pending.get(t, (torch.empty(0, dtype=torch.int32, device=self.device),) * 2)
            csr, keys = ops.build_csr(s, d, n)
            self._csr[t], self._keys[t] = csr, keys
        self._pending.clear()

    def csr(self, edge_type: str = "co_view") -> ops.CSRGraph:
        self.finalize()
        return self._csr[edge_type]

    def keys(self, edge_type: str) -> torch.Tensor:
        """Sorted-unique int64 keys src<<32|dst of one edge type."""
        self.finalize()
        return self._keys[edge_type]

    def neighbor_indices(self, i: int, edge_type: str = "co_view") -> torch.Tensor:
        g = self.csr(edge_type)
        lo, hi = g.rowptr[i: i + 2].tolist()
        return g.col[lo:hi].to(torch.int64)

    def similarity_keys(self) -> torch.Tensor:
        """(Bcv n Bpv) - Bcp  (synthetic_data.py:89, 118-119)."""
        return ops.set_difference(ops.set_intersection(self.keys("co_view"), self.keys("purchase_after_view")),
                                  self.keys("co_purchase"))

    def complementary_keys(self) -> torch.Tensor:
        """Bcp - (Bpv u Bcv)  (synthetic_data.py:90, 125-128)."""
        union = ops.set_union(self.keys("purchase_after_view"), self.keys("co_view"), self._num_nodes)
        return ops.set_difference(self.keys("co_purchase"), union)

    def derive_pair_sets(self) -> None:
        """Attach ``similarity_pairs`` / ``complementary_pairs`` as the reference generator does
        (synthetic_data.py:150-151); id tuples when the graph has string ids, else int tuples."""
        self.similarity_pairs = self._pairs_to_ids(self.similarity_keys())
        self.complementary_pairs = self._pairs_to_ids(self.complementary_keys())

    def _pairs_to_ids(self, keys: torch.Tensor) -> List[tuple]:
        s, d = ops.unpack_keys(keys)
        s, d = s.tolist(), d.tolist()
        if self._ids is None:
            return list(zip(s, d))
        return [(self._ids[a], self._ids[b]) for a, b in zip(s, d)]

    def type_members(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Type-sorted catalog permutation: (members int32 [P], offsets int64 [T+1]); members of type
        t are members[offsets[t]:offsets[t+1]] in ascending node order (get_products_by_type as a
        CSR, bpg.py:44-49).  Built with the same sort / CSR kernels as the edge lists."""
        self.finalize()
        if self._type_members is None:
            if self.type_id is None:
                raise ValueError("graph has no product types")
            n_types = int(self.type_id.max().item()) + 1 if self.type_id.numel() else 0
            node = torch.arange(self._num_nodes, dtype=torch.int32, device=self.device)
            csr, _ = ops.build_csr(self.type_id, node, n_types, self._num_nodes)
            self._type_members = (csr.col, csr.rowptr)
        return self._type_members
