"""Oracle (test infrastructure only): torch-CPU port of the reference modules.

The reference (/root/reference/src/models/*.py) is pure PyTorch; it cannot travel to the GPU
box, so this file restates it with the *same ATen ops* (nn.Linear / BatchNorm1d / Tanh /
nn.MultiheadAttention / F.pairwise_distance / torch.topk / torch.norm) and the same
``state_dict`` layout.  Uses: (1) ``bench.py``'s ``cpu_baseline`` and ``--impl reference``
legs time it on the host cores ("kind": "port"); (2) tests use its autograd as a second
gradient check.  ``tests/test_oracle_golden.py`` pins it against fixtures produced by the real
reference.  Never imported by the product package.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn


def default_config(**over) -> SimpleNamespace:
    """Duck-typed stand-in for /root/reference/config.py:5-57 without its mkdir side effects."""
    cfg = SimpleNamespace(PRODUCT_EMB_DIM=128, TYPE_EMB_DIM=64, HIDDEN_SIZE=256,
                          NUM_ATTENTION_HEADS=4, DROPOUT=0.1, MARGIN=1.0, NEG_SAMPLES=5,
                          BATCH_SIZE=256, LEARNING_RATE=1e-3, NUM_EPOCHS=20, ALPHA=0.8,
                          NUM_COMP_TYPES=3, NUM_TYPES=34800, PRODUCT2VEC_EPOCHS=10,
                          DEVICE=torch.device("cpu"))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


class PortProduct2Vec(nn.Module):
    """product2vec.py:8-81 (forward only; the training loop lives in bench/tests)."""

    def __init__(self, cfg):
        super().__init__()
        d, hid = cfg.PRODUCT_EMB_DIM, cfg.HIDDEN_SIZE
        layers = [nn.Linear(d, hid), nn.BatchNorm1d(hid), nn.Tanh(),
                  nn.Linear(hid, hid), nn.Tanh(), nn.Linear(hid, d)]
        self.ffn = nn.Sequential(*layers)
        self.attention = nn.MultiheadAttention(d, cfg.NUM_ATTENTION_HEADS,
                                               dropout=cfg.DROPOUT, batch_first=True)

    def embed(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() not in (1, 2, 3):
            raise ValueError(f"Unexpected input dimension: {x.dim()}")
        rows = x.reshape(-1, x.shape[-1])
        return self.ffn(rows).reshape(*x.shape[:-1], -1)

    def attend(self, q: torch.Tensor, kv: torch.Tensor) -> torch.Tensor:
        q3 = q.reshape(-1, 1, q.shape[-1])
        kv3 = kv if kv.dim() == 3 else kv.unsqueeze(0)
        out, _ = self.attention(q3, kv3, kv3)
        return out.reshape(q.shape)

    def forward(self, features: torch.Tensor, neighbors: Optional[torch.Tensor] = None):
        emb = self.embed(features)
        if neighbors is not None and neighbors.size(0) > 0:
            emb = self.attend(emb, self.embed(neighbors))
        return emb


def port_triplet_loss(model: PortProduct2Vec, batch: Dict[str, torch.Tensor], margin: float):
    """product2vec.py:132-154."""
    a = model(batch["anchor"], batch.get("anchor_neighbors"))
    p = model(batch["positive"])
    n = model(batch["negative"])
    dpos = F.pairwise_distance(a, p)
    dneg = F.pairwise_distance(a.unsqueeze(1).expand(-1, n.size(1), -1), n).mean(dim=1)
    return F.relu(margin - dpos + dneg).mean()


class PortPCompanion(nn.Module):
    """p_companion.py:9-119 on integer query indices."""

    def __init__(self, cfg, table: torch.Tensor):
        super().__init__()
        self.cfg = cfg
        L, d = cfg.TYPE_EMB_DIM, cfg.PRODUCT_EMB_DIM
        self.product_embeddings = nn.Embedding.from_pretrained(table, freeze=True)
        self.type_transition = nn.ModuleDict(dict(encoder=nn.Linear(L, L // 2), decoder=nn.Linear(L // 2, L)))
        self.item_prediction = nn.ModuleDict(dict(type_projection=nn.Linear(L, d), item_projection=nn.Linear(d, d)))
        self.query_type_embeddings = nn.Embedding(cfg.NUM_TYPES, L)
        self.complementary_type_embeddings = nn.Embedding(cfg.NUM_TYPES, L)
        self.drop = nn.Dropout(cfg.DROPOUT)

    def forward(self, query_idx, query_types):
        q = self.product_embeddings(query_idx)
        t = self.query_type_embeddings(query_types)
        base = self.type_transition["decoder"](self.drop(F.relu(self.type_transition["encoder"](t))))
        sims = base @ self.complementary_type_embeddings.weight.T
        top = torch.topk(sims, k=self.cfg.NUM_COMP_TYPES, dim=1).indices
        comp = self.complementary_type_embeddings(top)
        proj = self.item_prediction["item_projection"](q).unsqueeze(1) * self.item_prediction["type_projection"](comp)
        return {"projected_embeddings": proj, "complementary_types": top, "type_similarities": sims}

    def loss(self, out, pos_t, neg_t, pos_items, neg_items):
        r = torch.arange(out["type_similarities"].size(0))
        s = out["type_similarities"]
        tl = torch.clamp(self.cfg.MARGIN - s[r, pos_t] + s[r, neg_t], min=0).mean()
        pe = out["projected_embeddings"]
        dp = torch.norm(pe - pos_items.unsqueeze(1), dim=-1)
        dn = torch.norm(pe - neg_items.unsqueeze(1), dim=-1)
        il = torch.clamp(self.cfg.MARGIN - dp + dn, min=0).mean()
        return self.cfg.ALPHA * il + (1 - self.cfg.ALPHA) * tl


def port_masked_topk(q: torch.Tensor, catalog: torch.Tensor, row_type: torch.Tensor,
                     type_id: torch.Tensor, k: int):
    """inference.py:93-113 restated per score row: filter the catalog to the row's type,
    matmul, topk (fp32, torch's own tie order)."""
    outs, outi = [], []
    for r in range(q.shape[0]):
        sel = torch.nonzero(type_id == row_type[r]).squeeze(1)
        s = q[r:r + 1] @ catalog[sel].T
        kk = min(k, sel.numel())
        ts, ti = torch.topk(s[0], kk)
        outs.append(ts)
        outi.append(sel[ti])
    return outs, outi


def port_dense_topk(q: torch.Tensor, catalog: torch.Tensor, row_type: torch.Tensor,
                    type_id: torch.Tensor, k: int):
    """Dense whole-catalog form (north_star part 4): matmul + mask + topk in fp32."""
    s = q @ catalog.T
    s = s.masked_fill(type_id.unsqueeze(0) != row_type.unsqueeze(1), float("-inf"))
    return torch.topk(s, k, dim=1)
