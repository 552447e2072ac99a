"""P-Companion joint model on the B200 kernels.

Drop-ins for /root/reference/src/models/type_transition.py, item_prediction.py and
p_companion.py: same class names, constructor signatures, forward / compute_loss contracts,
attribute names and ``state_dict`` keys.  Type top-k (p_companion.py:64) and both hinge losses
(:95-119) run in hand-written kernels; ties in the type top-k rank the lower type index first.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import ops
from .dense import linear, type_scores


class ComplementaryTypeTransition(nn.Module):
    """decoder(dropout(relu(encoder(t)))), 64 -> 32 -> 64 (type_transition.py:5-20)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.encoder = nn.Linear(config.TYPE_EMB_DIM, config.TYPE_EMB_DIM // 2)
        self.decoder = nn.Linear(config.TYPE_EMB_DIM // 2, config.TYPE_EMB_DIM)
        self.dropout = nn.Dropout(config.DROPOUT)

    def forward(self, query_type_embedding):
        h = self.dropout(torch.relu(linear(query_type_embedding, self.encoder.weight, self.encoder.bias)))
        return linear(h, self.decoder.weight, self.decoder.bias)


class ComplementaryItemPrediction(nn.Module):
    """item_projection(q)[:, None, :] * type_projection(T) (item_prediction.py:5-39)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.type_projection = nn.Linear(config.TYPE_EMB_DIM, config.PRODUCT_EMB_DIM)
        self.item_projection = nn.Linear(config.PRODUCT_EMB_DIM, config.PRODUCT_EMB_DIM)

    def forward(self, query_item_embedding, complementary_type_embeddings):
        projected_item = linear(query_item_embedding, self.item_projection.weight, self.item_projection.bias)
        type_projections = linear(complementary_type_embeddings, self.type_projection.weight, self.type_projection.bias)
        return projected_item.unsqueeze(1) * type_projections


class PCompanion(nn.Module):
    """p_companion.py:9-119.  ``pretrained_embeddings`` may be the reference's Dict[str, Tensor]
    or a dense [P, D] tensor (then product ids are the integers 0..P-1 and ``batch['query_ids']``
    may be an index tensor - the path that avoids the per-step Python dict lookups of :47-49)."""

    def __init__(self, config, pretrained_embeddings):
        super().__init__()
        self.config = config
        if isinstance(pretrained_embeddings, torch.Tensor):
            embedding_matrix = pretrained_embeddings.detach().float()
            self.product_to_idx = None
        else:
            product_ids = list(pretrained_embeddings.keys())
            self.product_to_idx = {pid: idx for idx, pid in enumerate(product_ids)}
            embedding_matrix = torch.stack([pretrained_embeddings[p].detach().float().cpu() for p in product_ids])
        self.product_embeddings = nn.Embedding.from_pretrained(embedding_matrix, freeze=True)
        self.type_transition = ComplementaryTypeTransition(config)
        self.item_prediction = ComplementaryItemPrediction(config)
        self.query_type_embeddings = nn.Embedding(config.NUM_TYPES, config.TYPE_EMB_DIM)
        self.complementary_type_embeddings = nn.Embedding(config.NUM_TYPES, config.TYPE_EMB_DIM)

    def _query_indices(self, query_ids) -> torch.Tensor:
        dev = self.product_embeddings.weight.device
        if isinstance(query_ids, torch.Tensor):
            return query_ids.to(dev, torch.int64)
        if self.product_to_idx is None:
            return torch.as_tensor(list(query_ids), dtype=torch.int64, device=dev)
        return torch.tensor([self.product_to_idx[pid] for pid in query_ids], dtype=torch.int64, device=dev)  # KeyError as :48

    def forward(self, batch) -> Dict[str, torch.Tensor]:
        if not self.product_embeddings.weight.is_cuda:
            raise RuntimeError("PCompanion.forward: module is not on CUDA; pcompanion_b200 has no CPU fallback")
        query_embeddings = self.product_embeddings(self._query_indices(batch["query_ids"]))
        query_type_emb = self.query_type_embeddings(batch["query_types"])
        comp_base = self.type_transition(query_type_emb)
        similarities = type_scores(comp_base, self.complementary_type_embeddings.weight)     # [B, T]
        # the type loss reads two entries per row of this matrix; it takes its gradient path through the factors
        # (ops.type_hinge), so the dense [B, T] backward only runs if a caller differentiates the matrix itself
        similarities._pc_factors = (comp_base, self.complementary_type_embeddings.weight)
        _, top_types = ops.topk_rows(similarities.detach(), self.config.NUM_COMP_TYPES)
        comp_type_embeddings = self.complementary_type_embeddings(top_types)
        projected_embeddings = self.item_prediction(query_embeddings, comp_type_embeddings)
        return {
            "projected_embeddings": projected_embeddings,
            "complementary_types": top_types,
            "type_similarities": similarities,
        }

    def compute_loss(self, batch, outputs) -> torch.Tensor:
        type_loss = self._compute_type_loss(outputs["type_similarities"], batch["positive_types"].squeeze(-1),
                                            batch["negative_types"].squeeze(-1))
        item_loss = self._compute_item_loss(outputs["projected_embeddings"], batch["positive_items"],
                                            batch["negative_items"])
        return self.config.ALPHA * item_loss + (1 - self.config.ALPHA) * type_loss

    def _compute_type_loss(self, type_similarities, positive_types, negative_types):
        return ops.type_hinge(type_similarities, positive_types, negative_types, self.config.MARGIN)

    def _compute_item_loss(self, projected_embeddings, positive_items, negative_items):
        return ops.item_hinge(projected_embeddings, positive_items, negative_items, self.config.MARGIN)
