"""Top-k modules — drop-in for the reference's ``models/indexing/top_k.py``.

``MIPSBruteForceTopK.forward`` keeps the reference signature (top_k.py:44-70) but runs the
fused score + exact selection kernels (``grb_mips_topk``): the (B, X) logits never reach HBM.
Ties are broken by lowest item index (the reference's torch.topk order is arbitrary).
"""
from __future__ import annotations

import abc
from typing import Tuple

import torch

from . import functional as GF


class TopKModule(torch.nn.Module):
    @abc.abstractmethod
    def forward(self, query_embeddings: torch.Tensor, item_embeddings_t: torch.Tensor,
                item_ids: torch.Tensor, k: int, sorted: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (top_k_scores, top_k_ids), both (B, k)."""


class MIPSBruteForceTopK(TopKModule):
    def forward(self, query_embeddings: torch.Tensor, item_embeddings_t: torch.Tensor,
                item_ids: torch.Tensor, k: int, sorted: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """query_embeddings (B, D); item_embeddings_t (D, X) — the transposed view the
        reference's CandidateIndex keeps (candidate_index.py:29); item_ids (1, X) int64.
        The result is always sorted (``sorted=False`` only relaxes the reference's order)."""
        items = item_embeddings_t.t()  # (X, D); contiguous when the argument is the usual view
        scores, ids = GF.mips_topk(query_embeddings, items, item_ids.reshape(-1), k)
        if scores.dtype != query_embeddings.dtype and query_embeddings.dtype == torch.float32:
            scores = scores.to(query_embeddings.dtype)
        return scores, ids

    def forward_filtered(self, query_embeddings: torch.Tensor, item_embeddings_t: torch.Tensor,
                         item_ids: torch.Tensor, k: int, invalid_ids: torch.Tensor,
                         target_ids: torch.Tensor = None):
        """Top-k with the per-row invalid-id filter of candidate_index.py:125-158 (and, with
        ``target_ids``, the rank of metrics/retrieval.py:45-55) inside the selection kernel: the
        caller asks for k, not k + invalid_ids.size(1).  Returns (scores, ids[, ranks])."""
        items = item_embeddings_t.t()
        return GF.mips_topk(query_embeddings, items, item_ids.reshape(-1), k,
                            invalid_ids=invalid_ids, target_ids=target_ids)
