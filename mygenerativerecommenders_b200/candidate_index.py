"""Candidate index — drop-in for the reference's ``models/indexing/candidate_index.py``.

Same constructor, properties and ``get_top_k_outputs`` contract (candidate_index.py:9-164):
owns the id buffer (1, X) and the transposed view of the item table, widens k by the number
of invalid ids, asks the top-k module, drops invalid ids row-wise and returns (ids, scores).
``ShardedCandidateIndex`` adds the corpus-sharded variant: every rank holds
X/G items, runs the fused local top-k and the per-shard results are merged after an NCCL
all-gather with the same exact selection kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import functional as GF
from .top_k import TopKModule


def _drop_invalid(ids: torch.Tensor, scores: torch.Tensor, invalid_ids: torch.Tensor, k: int):
    """Keep, in order, the first k entries of each row whose id is not in invalid_ids[row]
    (candidate_index.py:142-158)."""
    bad = (ids.unsqueeze(2) == invalid_ids.unsqueeze(1)).any(dim=2)
    keep = ~bad
    keep &= torch.cumsum(keep.to(torch.int32), dim=1) <= k
    # stable argsort brings the kept positions to the front without reordering them
    pos = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :k]
    out_i, out_s = torch.gather(ids, 1, pos), torch.gather(scores, 1, pos)
    # rows left with fewer than k valid entries (the reference's .view(-1, k) raises there): pad like the
    # fused kernel does, (-1, -inf), instead of returning filtered-out ids
    short = ~torch.gather(keep, 1, pos)
    return out_i.masked_fill(short, -1), out_s.masked_fill(short, float("-inf"))


def _fused_filter_ok(top_k_module, k: int, n_invalid: int, num_objects: int) -> bool:
    """The selection kernel drops invalid ids itself when the list fits its shared memory and the
    reference's own result would hold k entries per row (k + n_invalid <= X)."""
    return (n_invalid > 0 and hasattr(top_k_module, "forward_filtered") and n_invalid <= 1024
            and k + n_invalid <= min(2048, num_objects))


class CandidateIndex(torch.nn.Module):
    def __init__(self, k: int, ids: torch.Tensor, top_k_module: TopKModule,
                 embeddings: torch.Tensor = None, invalid_ids: Optional[torch.Tensor] = None,
                 debug_path: Optional[str] = None) -> None:
        super().__init__()
        self.register_buffer("_ids", torch.as_tensor(ids).unsqueeze(0))
        self._k = min(k, self._ids.shape[1])
        self._top_k_module: TopKModule = top_k_module
        self._invalid_ids: Optional[torch.Tensor] = invalid_ids
        self._debug_path: Optional[str] = debug_path
        self.update_embeddings(embeddings)

    def update_embeddings(self, embeddings: Optional[torch.Tensor]) -> None:
        """embeddings: (1, X, D).  Stored as the (D, X) transposed *view* (no copy), like
        candidate_index.py:27-31."""
        self._embeddings_t = None if embeddings is None else embeddings.permute(2, 1, 0).squeeze(2)

    @property
    def ids(self) -> torch.Tensor:
        return self._ids

    @property
    def num_objects(self) -> int:
        return self._ids.size(1)

    @property
    def embeddings(self) -> Optional[torch.Tensor]:
        if self._embeddings_t is None:
            return None
        return self._embeddings_t.unsqueeze(2).permute(2, 1, 0).squeeze(2)

    def get_top_k_outputs(self, query_embeddings: torch.Tensor, k: int = None,
                          invalid_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (top_k_ids, top_k_scores), each (B, k) — ids first, as the reference."""
        n_invalid = invalid_ids.size(1) if invalid_ids is not None else 0
        if k is None:
            k = self._k
        if _fused_filter_ok(self._top_k_module, k, n_invalid, self.num_objects):
            scores, ids = self._top_k_module.forward_filtered(
                query_embeddings=query_embeddings, item_embeddings_t=self._embeddings_t,
                item_ids=self._ids, k=k, invalid_ids=invalid_ids)
            return ids, scores
        k_prime = min(k + n_invalid, self.num_objects)
        scores, ids = self._top_k_module(
            query_embeddings=query_embeddings, item_embeddings_t=self._embeddings_t,
            item_ids=self._ids, k=k_prime, sorted=True)
        if invalid_ids is not None:
            ids, scores = _drop_invalid(ids, scores, invalid_ids, k)
        return ids, scores

    def filter_invalid_ids(self, invalid_ids: torch.Tensor) -> "CandidateIndex":
        """candidate_index.py:52-105 builds a per-row copy of the whole index with the invalid ids removed
        (and, as written there, constructs ``CandidateIndex`` without its required ``k`` / ``top_k_module``
        arguments, so the reference's own method raises TypeError).  Per-row filtering is what
        ``get_top_k_outputs(..., invalid_ids=...)`` does inside the selection kernel."""
        raise NotImplementedError(
            "filter_invalid_ids: pass invalid_ids to get_top_k_outputs instead (filtered inside the top-k kernel)")

    def apply_object_filter(self) -> "CandidateIndex":
        raise NotImplementedError("not implemented.")


class ShardedCandidateIndex(CandidateIndex):
    """Corpus-sharded index: rank r of G holds the contiguous item range r*ceil(X/G) ...

    ``ids`` / ``embeddings`` passed to the constructor / ``update_embeddings`` are the FULL
    corpus (as in the reference, where every rank holds everything); this module keeps only
    its own slice.  Queries must be the same on every rank of ``group`` (replicated eval
    batch); every rank returns the full merged top-k."""

    def __init__(self, k: int, ids: torch.Tensor, top_k_module: TopKModule,
                 embeddings: torch.Tensor = None, invalid_ids: Optional[torch.Tensor] = None,
                 debug_path: Optional[str] = None, group=None) -> None:
        self._group = group
        self._world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._rank = dist.get_rank(group) if dist.is_initialized() else 0
        ids = torch.as_tensor(ids)
        self._num_total = ids.numel()
        self._k_total = min(k, self._num_total)
        lo, hi = self._range(self._num_total)
        super().__init__(k=k, ids=ids[lo:hi], top_k_module=top_k_module, embeddings=None,
                         invalid_ids=invalid_ids, debug_path=debug_path)
        self.update_embeddings(embeddings)

    def _range(self, n: int) -> Tuple[int, int]:
        per = -(-n // self._world)
        return min(self._rank * per, n), min((self._rank + 1) * per, n)

    def update_embeddings(self, embeddings: Optional[torch.Tensor]) -> None:
        if embeddings is not None and embeddings.size(1) == getattr(self, "_num_total", -1) \
                and self._world > 1:
            lo, hi = self._range(self._num_total)
            embeddings = embeddings[:, lo:hi].contiguous()
        super().update_embeddings(embeddings)

    @property
    def num_objects(self) -> int:
        return self._num_total

    def get_top_k_outputs(self, query_embeddings: torch.Tensor, k: int = None,
                          invalid_ids: Optional[torch.Tensor] = None):
        n_invalid = invalid_ids.size(1) if invalid_ids is not None else 0
        if k is None:
            k = self._k_total
        k_prime = min(k + n_invalid, self._num_total)
        fused = _fused_filter_ok(self._top_k_module, k, n_invalid, self._ids.size(1))
        if fused:
            # every shard filters its own result: the global top-k of valid items is contained in
            # the union of the shards' top-k of valid items, so k (not k') entries are exchanged
            k_prime, invalid_local, invalid_ids = k, invalid_ids, None
            scores, ids = self._top_k_module.forward_filtered(
                query_embeddings=query_embeddings, item_embeddings_t=self._embeddings_t,
                item_ids=self._ids, k=k, invalid_ids=invalid_local)
            k_local = k
        else:
            k_local = min(k_prime, self._ids.size(1))
            scores, ids = self._top_k_module(
                query_embeddings=query_embeddings, item_embeddings_t=self._embeddings_t,
                item_ids=self._ids, k=k_local, sorted=True)
        if self._world > 1:
            per = -(-self._num_total // self._world)
            k_locals = [min(k_prime, max(0, min((r + 1) * per, self._num_total) - min(r * per, self._num_total)))
                        for r in range(self._world)]
            scores, ids = merge_sharded_topk(scores.float(), ids, k_prime, self._world, self._group,
                                             k_locals=k_locals)
        if invalid_ids is not None:
            ids, scores = _drop_invalid(ids, scores, invalid_ids, k)
        return ids, scores


class _PeerExchange:
    """The exchange step of the sharded top-k over peer memory: every rank owns a symmetric
    (rows, world * k) gather buffer for scores and one for ids; ``put`` stores the local (B, k)
    block into its column block of every rank's buffer with peer-to-peer stores
    (``grb_p2p_put_rows``), a cross-rank barrier on the stream makes them visible, and the merge
    kernel reads the local buffer in place: no NCCL call, no host synchronisation, no transposing
    copy.  One instance per (group, rows, k); ``GRB_NO_P2P=1`` or an unavailable symmetric-memory
    backend falls back to the NCCL all-gather."""

    _cache: dict = {}
    _broken = False

    @classmethod
    def get(cls, group, rows: int, k: int, world: int, device):
        import os
        if cls._broken or os.environ.get("GRB_NO_P2P") == "1" or device.type != "cuda" \
                or dist.get_backend(group) != "nccl":
            return None
        key = (id(group), rows, k, world, device.index)
        st = cls._cache.get(key)
        if st is None:
            try:
                st = cls(group, rows, k, world, device)
            except Exception as e:   # no symmetric memory on this platform: keep the NCCL path
                cls._broken = True
                import warnings
                warnings.warn(f"grb200: peer-memory exchange unavailable ({e!r}); using NCCL all-gather")
                return None
            cls._cache[key] = st
        return st

    def __init__(self, group, rows: int, k: int, world: int, device) -> None:
        from .peer import PeerBarrier, symmetric_empty
        self.rows, self.k, self.world = rows, k, world
        self.rank = dist.get_rank(group)
        self.scores, self._hs, self.dst_scores = symmetric_empty((rows, world * k), torch.float32, device, group)
        self.ids, self._hi, self.dst_ids = symmetric_empty((rows, world * k), torch.int64, device, group)
        self.sync = PeerBarrier(group, device)

    def exchange(self, scores: torch.Tensor, ids: torch.Tensor):
        from . import _lib
        B, k = scores.shape
        scores, ids = scores.contiguous(), ids.contiguous()
        stream = _lib.stream_ptr(scores.device)
        self.sync.barrier(0, scores.device)   # every rank is done reading the previous round
        for src, dst, es in ((scores, self.dst_scores, 4), (ids, self.dst_ids, 8)):
            _lib.check(_lib.lib().grb_p2p_put_rows(
                src.data_ptr(), k * es, dst, self.world, self.world * self.k * es,
                self.rank * self.k * es, B, k * es, stream))
        self.sync.barrier(1, scores.device)   # all blocks have landed everywhere
        return self.scores[:B], self.ids[:B]


def merge_sharded_topk(scores: torch.Tensor, ids: torch.Tensor, k: int, world: int, group=None,
                       k_locals=None):
    """Exchange per-shard (B, k_local) results and select the exact global top-k.

    ``k_locals`` (every rank's k_local, known from the shard sizes without communication): when
    they are all equal the exchange runs over peer memory (``_PeerExchange``); otherwise — shards
    holding fewer than k items — short shards are padded with (-inf, int64 max) and the blocks go
    through an NCCL all-gather.  The selection is a kernel either way."""
    B, kl = scores.shape
    if k_locals is not None and len(set(k_locals)) == 1 and scores.is_cuda:
        px = _PeerExchange.get(group, max(B, 1), kl, world, scores.device)
        if px is not None:
            cand_s, cand_i = px.exchange(scores, ids)
            return GF.topk_merge(cand_s, cand_i, k)
    klen = torch.tensor([kl], device=scores.device, dtype=torch.int64)
    lens = [torch.zeros_like(klen) for _ in range(world)]
    dist.all_gather(lens, klen, group=group)
    kmax = int(max(int(x.item()) for x in lens))
    if kl < kmax:
        pad_s = torch.full((B, kmax - kl), float("-inf"), device=scores.device, dtype=scores.dtype)
        pad_i = torch.full((B, kmax - kl), torch.iinfo(torch.int64).max, device=ids.device, dtype=ids.dtype)
        scores, ids = torch.cat([scores, pad_s], 1), torch.cat([ids, pad_i], 1)
    # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
    all_s = torch.empty((world * B, kmax), device=scores.device, dtype=scores.dtype)
    all_i = torch.empty((world * B, kmax), device=ids.device, dtype=ids.dtype)
    dist.all_gather_into_tensor(all_s, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, ids.contiguous(), group=group)
    cand_s = all_s.view(world, B, kmax).permute(1, 0, 2).reshape(B, world * kmax)
    cand_i = all_i.view(world, B, kmax).permute(1, 0, 2).reshape(B, world * kmax)
    return GF.topk_merge(cand_s, cand_i, k)
