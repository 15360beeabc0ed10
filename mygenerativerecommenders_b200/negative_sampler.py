"""Negative samplers — drop-in for the reference's
``models/negatives_samples/negative_sampler.py`` (:21 base, :65 Local, :135 InBatch).

``forward(positive_ids, num_to_sample) -> (ids, embeddings)`` keeps the reference contract
(and its RNG stream: the same ``torch.randint`` call).  In addition every sampler exposes
``fused_sample`` which returns row indices into the source table(s) instead of the gathered
(N', R, D) tensor; :class:`~mygenerativerecommenders_b200.losses.SampledSoftmaxLoss` feeds
those to the fused gather.dot kernel, so the negatives tensor is never materialised.
"""
from __future__ import annotations

import abc
from typing import List, NamedTuple, Optional, Tuple

import torch


class FusedNegatives(NamedTuple):
    ids: torch.Tensor                 # (N', R) item ids compared against the positives
    table0: torch.Tensor              # (X0, d0) source table (autograd leaf or graph tensor)
    idx0: torch.Tensor                # (N', R) rows of table0
    table1: Optional[torch.Tensor]    # (X1, d1) second table, concatenated after table0
    idx1: Optional[torch.Tensor]
    l2_norm: bool                     # normalise the gathered rows inside the kernel
    zero_grad_rows: Tuple[Optional[int], Optional[int]]  # padding_idx rows of table0 / table1


def _reference_bases():
    """The reference's own sampler classes when its package is importable.  The reference trainer
    decides how to feed a sampler with ``isinstance(self.negatives_sampler, InBatchNegativesSampler)``
    against ITS class (models/retrieval.py:104); deriving from it keeps that check true when these
    classes are selected through the ``_target_`` override configs."""
    try:
        from generative_recommenders_pl.models.negatives_samples import negative_sampler as ref
        return ref.NegativesSampler, ref.LocalNegativesSampler, ref.InBatchNegativesSampler
    except Exception:       # no reference checkout on the path (the GPU box): plain modules
        return torch.nn.Module, None, None


_RefBase, _RefLocal, _RefInBatch = _reference_bases()


class NegativesSampler(_RefBase):
    def __init__(self, l2_norm: bool, l2_norm_eps: float) -> None:
        torch.nn.Module.__init__(self)      # not the reference's __init__ chain: same attributes below
        self._l2_norm: bool = l2_norm
        self._l2_norm_eps: float = l2_norm_eps

    def normalize_embeddings(self, x: torch.Tensor) -> torch.Tensor:
        return self._maybe_l2_norm(x)

    def _maybe_l2_norm(self, x: torch.Tensor) -> torch.Tensor:
        # x / clamp(||x||_2, min=eps)  (negative_sampler.py:31-37)
        if self._l2_norm:
            from . import functional as GF
            x = GF.l2_normalize(x, self._l2_norm_eps)
        return x

    @abc.abstractmethod
    def debug_str(self) -> str:
        pass

    @abc.abstractmethod
    def process_batch(self, ids: torch.Tensor, presences: torch.Tensor,
                      embeddings: torch.Tensor) -> None:
        pass

    @abc.abstractmethod
    def forward(self, positive_ids: torch.Tensor, num_to_sample: int) -> Tuple[torch.Tensor, torch.Tensor]:
        pass

    def fused_sample(self, positive_ids: torch.Tensor, num_to_sample: int) -> Optional[FusedNegatives]:
        """Row indices for the fused kernel, or None if this sampler cannot provide them."""
        return None


class LocalNegativesSampler(*((NegativesSampler, _RefLocal) if _RefLocal else (NegativesSampler,))):
    """Uniform sampling with replacement over ``all_item_ids`` (positives included)."""

    def __init__(self, l2_norm: bool, l2_norm_eps: float, num_items: int = None,
                 all_item_ids: List[int] = None) -> None:
        NegativesSampler.__init__(self, l2_norm=l2_norm, l2_norm_eps=l2_norm_eps)
        if all_item_ids is None and num_items is None:
            raise ValueError("Either num_items or all_item_ids must be provided")
        elif all_item_ids and num_items and num_items != len(all_item_ids):
            raise ValueError("num_items and all_item_ids must have the same length")
        elif all_item_ids:
            num_items = len(all_item_ids)
        elif num_items:
            all_item_ids = list(range(num_items))
        self._num_items: int = len(all_item_ids)
        self.register_buffer("_all_item_ids", torch.tensor(all_item_ids))
        self._item_emb: torch.nn.Embedding = None
        self._embeddings_module = None  # late-bound by the caller (retrieval.py:117)

    def debug_str(self) -> str:
        return f"local{f'-l2-eps{self._l2_norm_eps}' if self._l2_norm else ''}"

    def process_batch(self, ids, presences, embeddings) -> None:
        pass

    def _draw(self, positive_ids: torch.Tensor, num_to_sample: int) -> torch.Tensor:
        shape = positive_ids.size() + (num_to_sample,)
        offsets = torch.randint(low=0, high=self._num_items, size=shape,
                                dtype=positive_ids.dtype, device=positive_ids.device)
        return self._all_item_ids[offsets.view(-1)].reshape(shape)

    def forward(self, positive_ids: torch.Tensor, num_to_sample: int):
        sampled_ids = self._draw(positive_ids, num_to_sample)
        if self._embeddings_module is not None:
            emb = self._embeddings_module.get_item_embeddings(sampled_ids)
        else:
            emb = self._item_emb(sampled_ids)
        return sampled_ids, self.normalize_embeddings(emb)

    def fused_sample(self, positive_ids: torch.Tensor, num_to_sample: int) -> Optional[FusedNegatives]:
        mod = self._embeddings_module
        if mod is not None:
            item = getattr(mod, "_item_emb", None)
            year = getattr(mod, "_year_emb", None)
            if not isinstance(item, torch.nn.Embedding):
                return None
            if year is not None and not (isinstance(year, torch.nn.Embedding)
                                         and hasattr(mod, "lookup_year_ids")):
                return None
            # CategoricalEmbeddingModule-style id remapping is not a plain table gather
            if hasattr(mod, "_item_id_to_category_id"):
                return None
            ids = self._draw(positive_ids, num_to_sample)
            if year is None:
                return FusedNegatives(ids, item.weight, ids, None, None, self._l2_norm,
                                      (item.padding_idx, None))
            # this fork: concat(item_emb[id], year_emb[year_lookup[id]])  (embeddings.py:94-97)
            return FusedNegatives(ids, item.weight, ids, year.weight, mod.lookup_year_ids(ids),
                                  self._l2_norm, (item.padding_idx, year.padding_idx))
        if isinstance(self._item_emb, torch.nn.Embedding):
            ids = self._draw(positive_ids, num_to_sample)
            return FusedNegatives(ids, self._item_emb.weight, ids, None, None, self._l2_norm,
                                  (self._item_emb.padding_idx, None))
        return None


class InBatchNegativesSampler(*((NegativesSampler, _RefInBatch) if _RefInBatch else (NegativesSampler,))):
    """Uniform sampling over the ids present in the current batch (optionally de-duplicated)."""

    def __init__(self, l2_norm: bool, l2_norm_eps: float, dedup_embeddings: bool) -> None:
        NegativesSampler.__init__(self, l2_norm=l2_norm, l2_norm_eps=l2_norm_eps)
        self._dedup_embeddings: bool = dedup_embeddings

    def debug_str(self) -> str:
        s = f"in-batch{f'-l2-eps{self._l2_norm_eps}' if self._l2_norm else ''}"
        return s + "-dedup" if self._dedup_embeddings else s

    def process_batch(self, ids: torch.Tensor, presences: torch.Tensor,
                      embeddings: torch.Tensor) -> None:
        """ids (N') or (B, N) int64; presences same shape, bool; embeddings (..., D)."""
        assert ids.size() == presences.size()
        assert ids.size() == embeddings.size()[:-1]
        self._cache_valid(ids[presences], embeddings[presences])

    def _cache_valid(self, valid_ids: torch.Tensor, valid_emb: torch.Tensor) -> None:
        self._cached_count = None
        if self._dedup_embeddings:
            # one representative occurrence per distinct id (negative_sampler.py:168-184);
            # equal ids carry equal embeddings, so which occurrence wins does not matter
            uniq, inverse = torch.unique(valid_ids, sorted=False, return_inverse=True)
            rep = torch.empty(uniq.numel(), dtype=torch.int64, device=uniq.device)
            rep[inverse] = torch.arange(valid_ids.numel(), dtype=torch.int64, device=uniq.device)
            self._cached_embeddings = self._maybe_l2_norm(valid_emb[rep, :])
            self._cached_ids = uniq
        else:
            self._cached_embeddings = self._maybe_l2_norm(valid_emb)
            self._cached_ids = valid_ids

    def process_batch_prefix(self, ids: torch.Tensor, embeddings: torch.Tensor,
                             prefix_offsets: torch.Tensor, total: int,
                             static_shapes: bool = False, padded: bool = False) -> None:
        """``process_batch`` for the layout the training step has (generative_recommenders.py /
        retrieval.py:117-123): ids (B, N) whose non-zero entries are the first
        ``prefix_offsets[b+1] - prefix_offsets[b]`` of every row, ``total`` of them in all (known on
        the host from the batch lengths).  ``ids[presences]`` is then exactly the jagged packing
        of the rows, which needs no ``nonzero`` and therefore no device synchronisation; results
        are identical to ``process_batch(ids.view(-1), ids.view(-1) != 0, embeddings.view(-1, D))``."""
        from . import ops
        # padded: ``total`` is an upper bound (fixed row bucket); the real count is prefix_offsets[-1]
        valid_ids = ops.dense_to_jagged(ids.unsqueeze(-1), prefix_offsets, total=total,
                                        zero_tail=padded).squeeze(-1)
        valid_emb = ops.dense_to_jagged(embeddings, prefix_offsets, total=total, zero_tail=padded)
        if static_shapes and self._dedup_embeddings and total > 0:
            self._cache_valid_static(valid_ids, valid_emb, prefix_offsets[-1] if padded else None)
        else:
            self._cache_valid(valid_ids, valid_emb)

    def _cache_valid_static(self, valid_ids: torch.Tensor, valid_emb: torch.Tensor,
                            n_valid: Optional[torch.Tensor] = None) -> None:
        """De-duplication without data-dependent shapes, hence without a device sync: the cache is
        padded to the number of valid ids and the number of distinct ids stays on the device
        (``_cached_count``).  Same order as ``torch.unique`` (ascending ids).  ``_draw`` then maps
        random bits onto [0, count) on the device."""
        n = valid_ids.numel()
        dev = valid_ids.device
        pos = torch.arange(n, device=dev)
        live = pos < n_valid if n_valid is not None else None     # rows past n_valid are padding
        max_id = getattr(self, "max_item_id", None)
        if max_id is not None and max_id < (1 << 22):
            # small id space: direct addressing instead of a sort.  flags[id] = 1 for every id seen,
            # rank = exclusive prefix sum, the distinct ids in ascending order are the set positions
            key = valid_ids if live is None else torch.where(live, valid_ids, torch.zeros_like(valid_ids))
            flags = torch.zeros(max_id + 2, dtype=torch.int64, device=dev)
            flags.scatter_(0, key, torch.ones_like(key))
            flags[0:1].zero_()                                     # id 0 is padding, never a member
            rank_of_id = torch.cumsum(flags, 0) - flags            # exclusive prefix sum
            count = flags.sum()
            uniq = torch.nonzero_static(flags, size=n, fill_value=0).view(-1)
            # one representative row per distinct id; rows that are no members (padding) all go to
            # the last slot, which is free whenever such rows exist (count <= members < n)
            slot = torch.where(key != 0, rank_of_id[key], torch.full_like(key, n - 1))
            rep = torch.zeros(n, dtype=torch.int64, device=dev).scatter_(0, slot, pos)
        else:
            if live is not None:
                valid_ids = torch.where(live, valid_ids,
                                        torch.full_like(valid_ids, torch.iinfo(torch.int64).max))
            sorted_ids, order = torch.sort(valid_ids)
            new = torch.ones(n, dtype=torch.bool, device=dev)
            new[1:] = sorted_ids[1:] != sorted_ids[:-1]
            rank = torch.cumsum(new, 0) - 1
            uniq = torch.zeros(n, dtype=torch.int64, device=dev).scatter_(0, rank, sorted_ids)
            # (slots past the count keep distinct positions; nothing samples them, their gradient is 0)
            rep = pos.clone().scatter_(0, rank, order)
            if n_valid is None:
                count = rank[-1] + 1
            else:
                # (index_select, not rank[i]: indexing with a 0-dim tensor reads it on the host)
                count = rank.index_select(0, (n_valid - 1).clamp(min=0).view(1)).view(()) + 1
        from . import functional as GF
        self._cached_embeddings = self._maybe_l2_norm(GF.embedding_lookup(valid_emb, rep, None))
        self._cached_ids = uniq
        self._cached_count = count

    def process_batch_table(self, ids: torch.Tensor, prefix_offsets: torch.Tensor, total: int,
                            table: torch.Tensor, padded: bool = False, grad_scope=None,
                            rows_extra: int = 0) -> bool:
        """``process_batch`` straight from the embedding table (extension): equal ids carry equal
        embeddings, so the de-duplicated cache is ``normalize(table[unique(valid ids)])`` -- the
        (B, N, D) embeddings of the batch (retrieval.py:104-111 looks them up a second time) are not
        needed at all.  ids (B, N) whose valid entries are the first prefix_offsets[b+1] -
        prefix_offsets[b] of every row, ``total`` of them (an upper bound when ``padded``).  Static
        shapes, no device sync.  Needs ``dedup_embeddings`` and a small id space (``max_item_id``);
        returns False (nothing done) otherwise.  ``rows_extra``: the valid count of row b is
        prefix_offsets[b+1] - prefix_offsets[b] + rows_extra (the caller's jagged offsets + the target)."""
        from . import functional as GF
        from . import ops
        max_id = getattr(self, "max_item_id", None)
        if not self._dedup_embeddings or max_id is None or max_id >= (1 << 22) or total <= 0:
            return False
        if ids.is_cuda and ids.dtype == torch.int64 and ids.dim() == 2 and ids.is_contiguous() \
                and prefix_offsets.dtype in (torch.int32, torch.int64) and ids.size(0) <= 65535:
            # membership table stamped with a device-side epoch + one compaction: two launches
            # (grb_inbatch_distinct_ids) instead of the twelve of the torch formulation below
            uniq, count = self._distinct_ids(ids, prefix_offsets.contiguous(), int(total), int(rows_extra), max_id + 2)
            self._cached_embeddings = self._maybe_l2_norm(GF.embedding_lookup(table, uniq, 0, grad_scope=grad_scope))
            self._cached_ids = uniq
            self._cached_count = count
            return True
        if rows_extra:
            prefix_offsets = prefix_offsets + rows_extra * torch.arange(
                prefix_offsets.numel(), device=prefix_offsets.device, dtype=prefix_offsets.dtype)
        valid_ids = ops.dense_to_jagged(ids.unsqueeze(-1), prefix_offsets, total=total,
                                        zero_tail=padded).squeeze(-1)
        n = valid_ids.numel()
        dev = valid_ids.device
        if padded:       # rows past the real count are padding: id 0
            live = torch.arange(n, device=dev) < prefix_offsets[-1]
            valid_ids = torch.where(live, valid_ids, torch.zeros_like(valid_ids))
        flags = torch.zeros(max_id + 2, dtype=torch.int64, device=dev)
        flags.scatter_(0, valid_ids, torch.ones_like(valid_ids))
        flags[0:1].zero_()                                     # id 0 is padding, never a member
        count = flags.sum()
        uniq = torch.nonzero_static(flags, size=n, fill_value=0).view(-1)   # ascending ids, then zeros
        # slots past the count hold id 0 = the padding row: zero embedding, no gradient, never sampled
        self._cached_embeddings = self._maybe_l2_norm(GF.embedding_lookup(table, uniq, 0, grad_scope=grad_scope))
        self._cached_ids = uniq
        self._cached_count = count
        return True

    def _distinct_ids(self, ids: torch.Tensor, offsets: torch.Tensor, n_out: int, rows_extra: int,
                      n_flags: int) -> Tuple[torch.Tensor, torch.Tensor]:
        from . import _lib
        dev = ids.device
        tab = self.__dict__.get("_flag_table")
        n_tab = n_flags + (n_flags + 1023) // 1024      # membership flags + one counter per 1024 ids
        if tab is None or tab[0].device != dev or tab[0].numel() != n_tab:
            # persistent: zero-filled once, entries are stamped with the epoch, never cleared
            tab = (torch.zeros(n_tab, dtype=torch.int32, device=dev),
                   torch.ones(1, dtype=torch.int32, device=dev))
            self.__dict__["_flag_table"] = tab
        flags, epoch = tab
        uniq = torch.empty(n_out, dtype=torch.int64, device=dev)
        count = torch.empty(1, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().grb_inbatch_distinct_ids(
            ids.data_ptr(), ids.size(0), ids.size(1), offsets.data_ptr(), _lib.index_bits(offsets), rows_extra,
            flags.data_ptr(), n_flags, epoch.data_ptr(), uniq.data_ptr(), n_out, count.data_ptr(),
            _lib.stream_ptr(dev)))
        return uniq, count[0]

    def get_all_ids_and_embeddings(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if getattr(self, "_cached_count", None) is not None:     # padded cache: trim (host sync)
            c = int(self._cached_count.item())
            return self._cached_ids[:c], self._cached_embeddings[:c]
        return self._cached_ids, self._cached_embeddings

    def _draw(self, positive_ids: torch.Tensor, num_to_sample: int) -> torch.Tensor:
        size = positive_ids.size() + (num_to_sample,)
        if getattr(self, "_cached_count", None) is not None:
            # uniform over [0, count) without reading count on the host: 62 random bits modulo
            # count (bias < count / 2^62).  Same distribution as torch.randint(0, count), not the
            # same Philox stream (ATen folds the range into the draw on the host side).
            raw = torch.randint(low=0, high=2 ** 62, size=size, dtype=torch.int64,
                                device=positive_ids.device)
            return (raw % self._cached_count.clamp(min=1)).to(positive_ids.dtype)
        return torch.randint(low=0, high=self._cached_ids.size(0), size=size,
                             dtype=positive_ids.dtype, device=positive_ids.device)

    def forward(self, positive_ids: torch.Tensor, num_to_sample: int):
        offsets = self._draw(positive_ids, num_to_sample)
        return self._cached_ids[offsets], self._cached_embeddings[offsets]

    def _draw_with_ids(self, positive_ids: torch.Tensor, num_to_sample: int):
        """(offsets, ids) of the static-shape cache in ONE kernel (``grb_draw_negatives``: Philox draw,
        modulo the device-side count, id gather) instead of randint + remainder + index; falls back to
        ``_draw`` + a gather when ``_draw`` was replaced (tests inject fixed draws) or the cache is the
        ``unique`` one."""
        from . import _lib
        if (getattr(self, "_cached_count", None) is None or "_draw" in self.__dict__
                or not positive_ids.is_cuda or self._cached_ids.dtype != torch.int64):
            offsets = self._draw(positive_ids, num_to_sample)
            return offsets, self._cached_ids[offsets]
        dev = positive_ids.device
        size = tuple(positive_ids.size()) + (num_to_sample,)
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=dev)     # torch's generator: graph safe
        offsets = torch.empty(size, dtype=torch.int64, device=dev)
        ids = torch.empty(size, dtype=torch.int64, device=dev)
        count = self._cached_count.reshape(1)
        if count.dtype != torch.int64:
            count = count.to(torch.int64)
        _lib.check(_lib.lib().grb_draw_negatives(seed.data_ptr(), count.data_ptr(), self._cached_ids.data_ptr(),
                                                 offsets.numel(), offsets.data_ptr(), ids.data_ptr(),
                                                 _lib.stream_ptr(dev)))
        return offsets, ids

    def fused_sample(self, positive_ids: torch.Tensor, num_to_sample: int) -> Optional[FusedNegatives]:
        offsets, ids = self._draw_with_ids(positive_ids, num_to_sample)
        # the cache is already normalised (process_batch), so the kernel must not re-normalise
        return FusedNegatives(ids, self._cached_embeddings, offsets, None, None, False, (None, None))
