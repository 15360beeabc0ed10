"""Retrieval metrics — drop-in for the reference's ``models/metrics/retrieval.py`` without torchmetrics.

``RetrievalMetrics.compute`` (metrics/retrieval.py:40-68) derives every figure from ``ranks``: 1 + the
position of the target id in the row's top-k ids, k + 1 when it is absent.  The fused top-k kernel
returns exactly that per query (``MIPSBruteForceTopK.forward_filtered(..., target_ids=...)``), so the
evaluation loop can hand over ranks (``update_ranks``) instead of the (B, k) id matrix; ``update`` keeps the
reference's ``(top_k_ids, target_ids)`` signature.  States are plain tensors kept on the device of the
first update; ``compute`` returns the reference's dict: ``ndcg@K`` and ``hr@K`` for every K in
``at_k_list``, and ``mrr``.  (Multi-process aggregation: all-gather ``ranks()`` and feed ``update_ranks``.)
"""
from __future__ import annotations

from typing import Dict, List

import torch


class RetrievalMetrics:
    def __init__(self, k: int, at_k_list: List[int], **kwargs) -> None:
        self.k = k
        self.at_k_list = list(at_k_list)
        self._ranks: List[torch.Tensor] = []

    def reset(self) -> None:
        self._ranks = []

    def update(self, top_k_ids: torch.Tensor, target_ids: torch.Tensor, **kwargs) -> None:
        """top_k_ids (B, k), target_ids (B, 1) or (B,): metrics/retrieval.py:36-38 + :45-54."""
        assert top_k_ids.size(1) == self.k
        tgt = target_ids.reshape(-1, 1)
        _, idx = torch.max(torch.cat([top_k_ids, tgt], dim=1) == tgt, dim=1)
        self._ranks.append(idx + 1)

    def update_ranks(self, ranks: torch.Tensor) -> None:
        """ranks (B,): 1 + position of the target in the row's top-k, k + 1 when absent (the third output
        of the fused top-k with ``target_ids``)."""
        self._ranks.append(ranks.reshape(-1).to(torch.int64))

    def ranks(self) -> torch.Tensor:
        return torch.cat(self._ranks) if self._ranks else torch.zeros(0, dtype=torch.int64)

    def compute(self) -> Dict[str, torch.Tensor]:
        ranks = self.ranks()
        out: Dict[str, torch.Tensor] = {}
        zero = torch.zeros(1, dtype=torch.float32, device=ranks.device)
        for at_k in self.at_k_list:            # metrics/retrieval.py:57-62
            out[f"ndcg@{at_k}"] = torch.where(ranks <= at_k, 1.0 / torch.log2(ranks + 1), zero).mean()
        for at_k in self.at_k_list:            # :64-65
            out[f"hr@{at_k}"] = (ranks <= at_k).to(torch.float32).mean()
        out["mrr"] = (1.0 / ranks).mean()      # :67
        return out
