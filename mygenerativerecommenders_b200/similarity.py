"""Similarity modules — drop-in for the reference's ``models/similarity/{ndp_module,
dot_product}.py``.  The three shape branches and their (inconsistent) return types are kept
exactly: branch 1 returns ``(logits, {})``, branches 2/3 a bare tensor (dot_product.py:49-64).
These are plain library GEMMs; the fused paths (top-k, sampled softmax) bypass them."""
from __future__ import annotations

from typing import Optional

import torch


class NDPModule(torch.nn.Module):
    def forward(self, input_embeddings: torch.Tensor, item_embeddings: torch.Tensor,
                item_sideinfo: Optional[torch.Tensor], item_ids: torch.Tensor,
                precomputed_logits: Optional[torch.Tensor] = None) -> torch.Tensor:
        pass


class DotProductSimilarity(NDPModule):
    def __init__(self) -> None:
        super().__init__()

    def debug_str(self) -> str:
        return "dp"

    def forward(self, input_embeddings: torch.Tensor, item_embeddings: torch.Tensor,
                item_sideinfo: Optional[torch.Tensor], item_ids: torch.Tensor,
                precomputed_logits: Optional[torch.Tensor] = None):
        del item_ids
        if item_embeddings.size(0) == 1:
            # (B, D) x (1, X, D) -> (B, X), returned as a tuple (dot_product.py:49-54)
            return torch.mm(input_embeddings, item_embeddings.squeeze(0).t()), {}
        if input_embeddings.size(0) != item_embeddings.size(0):
            # (B*r, D) x (B, X, D) -> (B*r, X)
            B, X, D = item_embeddings.size()
            return torch.bmm(input_embeddings.view(B, -1, D),
                             item_embeddings.transpose(1, 2)).view(-1, X)
        # (B, D) x (B, X, D) -> (B, X)
        return torch.bmm(item_embeddings, input_embeddings.unsqueeze(2)).squeeze(2)
