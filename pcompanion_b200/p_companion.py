"""P-Companion joint model on the B200 kernels.

Drop-ins for /root/reference/src/models/type_transition.py, item_prediction.py and
p_companion.py: same class names, constructor signatures, forward / compute_loss contracts,
attribute names and ``state_dict`` keys.  Every layer runs in hand-written kernels: the type-transition MLP with its embedding gather
(csrc/pcomp.cu), the [B, L] x [L, T] type scoring with the top-k in the GEMM epilogue and the two projections of the item
prediction (tcgen05, csrc/gemm.cu), both hinge losses (:95-119, csrc/loss.cu); gradients of the embedding tables are
deterministic segment sums.  Ties in the type top-k rank the lower type index first.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import ops
from .dense import linear, type_scores_topk


def _dropout_args(module: nn.Module, p: float):
    """(p, seed) of the counter-based dropout mask of this call: the seed is drawn from torch's CPU generator, so
    torch.manual_seed makes runs reproducible; (0, 0) in eval mode."""
    if module.training and p > 0.0:
        return float(p), int(torch.randint(0, 2 ** 62, (1,)).item())
    return 0.0, 0


class ComplementaryTypeTransition(nn.Module):
    """decoder(dropout(relu(encoder(t)))), 64 -> 32 -> 64 (type_transition.py:5-20) as one fused kernel each way
    (csrc/pcomp.cu).  ``forward_indexed`` takes the type table and the indices instead of the gathered rows."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.encoder = nn.Linear(config.TYPE_EMB_DIM, config.TYPE_EMB_DIM // 2)
        self.decoder = nn.Linear(config.TYPE_EMB_DIM // 2, config.TYPE_EMB_DIM)
        self.dropout = nn.Dropout(config.DROPOUT)

    def _native(self) -> bool:
        return ops.mlp2_supported(self.encoder.in_features, self.encoder.out_features, self.decoder.out_features)

    def _run(self, table, index):
        p, seed = _dropout_args(self, self.dropout.p)
        counter = getattr(self, "_replay_counter", None)       # set by graphs.GraphedTrainStep: fresh mask per graph replay
        out = ops.mlp2(table, index, self.encoder.weight, self.encoder.bias, self.decoder.weight, self.decoder.bias, p, seed,
                       counter if p > 0.0 else None)
        if counter is not None and p > 0.0:
            counter.add_(1)
        return out

    def forward(self, query_type_embedding):
        if not query_type_embedding.is_cuda:
            raise RuntimeError("ComplementaryTypeTransition.forward: input is not on CUDA; pcompanion_b200 has no CPU fallback")
        if self._native() and query_type_embedding.dtype == torch.float32:
            lead = query_type_embedding.shape[:-1]
            out = self._run(query_type_embedding.reshape(-1, query_type_embedding.shape[-1]), None)
            return out.reshape(*lead, out.shape[-1])
        h = self.dropout(torch.relu(linear(query_type_embedding, self.encoder.weight, self.encoder.bias)))
        return linear(h, self.decoder.weight, self.decoder.bias)

    def forward_indexed(self, type_table: torch.Tensor, type_index: torch.Tensor):
        """forward(type_table[type_index]) without materialising the gathered rows (p_companion.py:54-57)."""
        if self._native() and type_table.dtype == torch.float32:
            return self._run(type_table, type_index)
        return self.forward(ops.gather_rows(type_table, type_index))


class ComplementaryItemPrediction(nn.Module):
    """item_projection(q)[:, None, :] * type_projection(T) (item_prediction.py:5-39): two tcgen05 projections and one
    broadcast-multiply kernel."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.type_projection = nn.Linear(config.TYPE_EMB_DIM, config.PRODUCT_EMB_DIM)
        self.item_projection = nn.Linear(config.PRODUCT_EMB_DIM, config.PRODUCT_EMB_DIM)

    def forward(self, query_item_embedding, complementary_type_embeddings):
        if not query_item_embedding.is_cuda:
            raise RuntimeError("ComplementaryItemPrediction.forward: input is not on CUDA; pcompanion_b200 has no CPU fallback")
        b, kt, l = complementary_type_embeddings.shape
        return self.forward_rows(query_item_embedding, complementary_type_embeddings.reshape(b * kt, l), kt)

    def forward_rows(self, query_item_embedding, type_rows, kt: int):
        """type_rows [B * Kt, L]: the complementary type embeddings of row b at rows b * Kt .. b * Kt + Kt - 1."""
        projected_item = linear(query_item_embedding, self.item_projection.weight, self.item_projection.bias)
        type_projections = linear(type_rows, self.type_projection.weight, self.type_projection.bias)
        if projected_item.dtype == torch.float32 and projected_item.shape[1] % 4 == 0:
            return ops.item_combine(projected_item, type_projections, kt)
        return projected_item.unsqueeze(1) * type_projections.reshape(projected_item.shape[0], kt, -1)


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
        kt = self.config.NUM_COMP_TYPES
        dev = self.product_embeddings.weight.device
        query_embeddings = ops.gather_rows(self.product_embeddings.weight, self._query_indices(batch["query_ids"]))
        query_types = batch["query_types"].to(dev)
        comp_base = self.type_transition.forward_indexed(self.query_type_embeddings.weight, query_types)
        # [B, L] x [L, T] on the tensor cores with the row top-Kt kept in the epilogue; the [B, T] matrix is written once
        # because it is part of forward()'s contract.  The type loss reads two entries per row of it and takes its
        # gradient path through the factors (ops.type_hinge), so no dense [B, T] backward runs unless a caller
        # differentiates the matrix itself.
        similarities, top_types = type_scores_topk(comp_base, self.complementary_type_embeddings.weight, kt)
        similarities._pc_factors = (comp_base, self.complementary_type_embeddings.weight)
        comp_rows = ops.gather_rows(self.complementary_type_embeddings.weight, top_types)            # [B * Kt, L]
        projected_embeddings = self.item_prediction.forward_rows(query_embeddings, comp_rows, kt)
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
