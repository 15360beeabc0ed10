"""Lightning-free host harness around the two hot paths: what the reference's ``Retrieval``
LightningModule does in ``training_step`` / ``retrieve`` (models/retrieval.py:21-48, :80-146,
:162-169; models/generative_recommenders.py:355-425), minus Lightning/Hydra, with this package's
modules wired in where the reference's ``_target_`` strings point (configs/model/hstu.yaml).

The pieces either side of the hot paths — embedding tables, the positional pre-processor, the
L2 post-processor and ``seq_features_from_row`` — are the reference's callers; they are restated
here on plain torch ops so bench.py and the tests can drive a whole step on a box that has no
reference tree.  They are SURVEY §8(f) "next" rows, not B200 kernels yet.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, NamedTuple, Optional, Tuple

import torch

from . import functional as GF
from . import ops
from .candidate_index import CandidateIndex, ShardedCandidateIndex
from .hstu import HSTU
from .losses import SampledSoftmaxLoss
from .negative_sampler import InBatchNegativesSampler, LocalNegativesSampler
from .similarity import DotProductSimilarity
from .top_k import MIPSBruteForceTopK


@dataclass
class RetrievalConfig:
    """One row of SURVEY §8's config table."""
    name: str = "ml-1m"
    num_items: int = 3952            # embedding rows = num_items + 1 (id 0 = padding)
    max_sequence_length: int = 200
    gr_output_length: int = 10
    embedding_dim: int = 50
    num_blocks: int = 2
    num_heads: int = 1
    attention_dim: int = 50
    linear_dim: int = 50
    dropout: float = 0.2
    sampler: str = "local"           # "local" | "inbatch"
    num_negatives: int = 128
    temperature: float = 0.05
    l2_eps: float = 1e-6
    top_k: int = 200
    split_year_embedding: bool = False  # this fork's item+year concat (embeddings.py:55-97)
    compute_dtype: Optional[torch.dtype] = None

    @property
    def N(self) -> int:
        return self.max_sequence_length + self.gr_output_length + 1


class SequentialFeatures(NamedTuple):
    past_lengths: torch.Tensor
    past_ids: torch.Tensor
    past_embeddings: Optional[torch.Tensor]
    past_payloads: Dict[str, torch.Tensor]


def seq_features_from_row(row: Dict[str, torch.Tensor], device, max_output_length: int):
    """models/utils/features.py:19-85: move to device, right-pad by ``max_output_length`` zeros,
    write the target timestamp at index ``length``."""
    lengths = row["history_lengths"].to(device, non_blocking=True)
    ids = row["historical_ids"].to(device, non_blocking=True)
    ts = row["historical_timestamps"].to(device, non_blocking=True)
    target_ids = row["target_ids"].to(device, non_blocking=True).unsqueeze(1)
    target_ts = row["target_timestamps"].to(device, non_blocking=True).unsqueeze(1)
    return _features_on_device(lengths, ids, ts, target_ids, target_ts, max_output_length)


def _features_on_device(lengths, ids, ts, target_ids, target_ts, max_output_length: int):
    if max_output_length > 0:
        ids = torch.nn.functional.pad(ids, (0, max_output_length))
        ts = torch.nn.functional.pad(ts, (0, max_output_length))
        ts.scatter_(dim=1, index=lengths.view(-1, 1), src=target_ts.view(-1, 1))
    return SequentialFeatures(lengths, ids, None, {"timestamps": ts}), target_ids


class ItemEmbeddings(torch.nn.Module):
    """models/embeddings/embeddings.py:40-101.  ``split_year``: concat(item_emb[id],
    year_emb[year_lookup[id]]), each D/2 wide, as in this fork; else one D-wide table."""

    def __init__(self, num_items: int, dim: int, split_year: bool) -> None:
        super().__init__()
        self._item_embedding_dim = dim
        if split_year:
            self._item_emb = torch.nn.Embedding(num_items + 1, dim // 2, padding_idx=0)
            self._year_emb = torch.nn.Embedding(num_items + 1, dim - dim // 2, padding_idx=0)
            self.register_buffer("year_lookup_table", torch.zeros(num_items + 1, dtype=torch.long))
        else:
            self._item_emb = torch.nn.Embedding(num_items + 1, dim, padding_idx=0)
            self._year_emb = None
        for p in self.parameters():
            torch.nn.init.trunc_normal_(p, mean=0.0, std=0.02, a=-0.04, b=0.04)

    def lookup_year_ids(self, item_ids: torch.Tensor) -> torch.Tensor:
        return self.year_lookup_table[item_ids.clamp(0, self.year_lookup_table.size(0) - 1)]

    def get_item_embeddings(self, item_ids: torch.Tensor) -> torch.Tensor:
        look = lambda emb, ids: GF.embedding_lookup(emb.weight, ids, emb.padding_idx)   # CUDA only
        if self._year_emb is None:
            return look(self._item_emb, item_ids)
        return torch.cat([look(self._item_emb, item_ids),
                          look(self._year_emb, self.lookup_year_ids(item_ids))], dim=-1)


class PositionalPreprocessor(torch.nn.Module):
    """preprocessors/learnable_positional_embedding.py:42-58: emb*sqrt(D) + pos_emb, dropout,
    zero the padded positions."""

    def __init__(self, max_sequence_len: int, dim: int, dropout: float) -> None:
        super().__init__()
        self._embedding_dim = dim
        self._pos_emb = torch.nn.Embedding(max_sequence_len, dim)
        torch.nn.init.trunc_normal_(self._pos_emb.weight, std=math.sqrt(1.0 / dim),
                                    a=-2 * math.sqrt(1.0 / dim), b=2 * math.sqrt(1.0 / dim))
        self._emb_dropout = torch.nn.Dropout(p=dropout)

    def forward(self, past_lengths, past_ids, past_embeddings, past_payloads):
        N = past_ids.size(1)
        x = past_embeddings * (self._embedding_dim ** 0.5) + self._pos_emb.weight[:N].unsqueeze(0)
        x = self._emb_dropout(x)
        valid = (past_ids != 0).unsqueeze(-1).float()
        return past_lengths, x * valid, valid, None


class L2NormPostprocessor(torch.nn.Module):
    """postprocessors/postprocessors.py:47-55."""

    def __init__(self, dim: int, eps: float = 1e-6) -> None:
        super().__init__()
        self._embedding_dim, self._eps = dim, eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x[..., : self._embedding_dim]
        return GF.l2_normalize(x, self._eps)


class PeerTableGrads:
    """Data-parallel exchange of the item-table gradient over peer memory instead of a dense
    all-reduce (the reference all-reduces the whole (V, D) table under Lightning DDP; at the ml-20m
    shape that is 134 MB per step although a batch touches ~11 k rows).

    Every rank owns TWO symmetric staging buffers (used by alternate steps), each with one (cap, D)
    block of rows + ids per source rank.  After backward each rank scales the rows its batch touched by
    1 / world in place in its own dense gradient (``grb_rows_scale``), stores them into its block on
    every OTHER rank (``grb_p2p_put_table_rows``: plain 16-byte stores over NVLink / NVSwitch — remote
    atomics are several times slower), one stream barrier, then adds the peers' rows to its dense
    gradient with local atomics (``grb_rows_scatter_add``): the tensor autograd produced becomes the
    averaged gradient, no second (V, D) buffer, no zero fill.  The double buffer makes the second
    barrier per step unnecessary: a rank that is one step ahead writes the OTHER buffer, and it cannot
    be two steps ahead because of this step's barrier.  Traffic per rank: touched_rows x D x 4 bytes per
    peer.  Exclude the parameter from DDP — this object does its reduction."""

    def __init__(self, weight: torch.nn.Parameter, group=None, hook: bool = True, sync=None) -> None:
        import torch.distributed as dist
        from .peer import PeerBarrier
        self.weight = weight
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.sync = sync if sync is not None else PeerBarrier(group, weight.device)
        self.cap = 0                       # staging rows per source rank, fixed by the first batch
        self.step = 0
        self.touched: Optional[torch.Tensor] = None
        dist.broadcast(weight.data, src=dist.get_global_rank(self.group, 0), group=group)
        if hook:                           # stand-alone use next to DDP: reduce as soon as the gradient exists
            weight.register_post_accumulate_grad_hook(self._hook)

    @torch.no_grad()
    def note_ids(self, *id_tensors: torch.Tensor) -> None:
        """Record the table rows this step's batch reads (distinct ids, padded with 0 = skipped).
        Static shapes, no host synchronisation."""
        V = self.weight.shape[0]
        flags = torch.zeros(V, dtype=torch.bool, device=self.weight.device)
        n = 0
        for t in id_tensors:
            flags.index_fill_(0, t.reshape(-1), True)
            n += t.numel()
        flags[0:1].zero_()
        self.touched = torch.nonzero_static(flags, size=n, fill_value=0).view(-1)

    def _allocate(self, n: int, D: int, device) -> None:
        import torch.distributed as dist
        from .peer import pointer_array, symmetric_empty
        self.cap = n
        self.stage = []
        for _ in range(2):
            rows, hr, _ = symmetric_empty((self.world * n, D), torch.float32, device, self.group)
            ids, hi, _ = symmetric_empty((self.world * n,), torch.int64, device, self.group)
            ids.zero_()                    # the own block is never written: id 0 = skipped by the scatter
            peers = [r for r in range(self.world) if r != self.rank]
            self.stage.append(dict(
                rows=rows, ids=ids, handles=(hr, hi),
                dst_rows=pointer_array([hr.buffer_ptrs[r] for r in peers]),
                dst_ids=pointer_array([hi.buffer_ptrs[r] for r in peers])))
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)     # nobody stores into a peer's buffer before it is zeroed

    def _hook(self, param: torch.nn.Parameter) -> None:
        self.put(param)
        self.sync.barrier(0, param.device)     # every rank's block has landed everywhere
        self.add_peers(param)

    def put(self, param: torch.nn.Parameter) -> None:
        """Scale the touched rows of the local gradient in place and store them into every peer's
        staging buffer (stream ordered; needs a cross-rank barrier before ``add_peers``)."""
        from . import _lib
        if self.touched is None or param.grad is None:
            raise RuntimeError("PeerTableGrads: note_ids() was not called for this step")
        g = param.grad
        if g.dtype != torch.float32 or not g.is_contiguous():
            g = g.float().contiguous()
            param.grad = g
        V, D = g.shape
        n = self.touched.numel()
        if self.cap == 0:
            self._allocate(n, D, g.device)
        elif n != self.cap:
            raise RuntimeError(f"PeerTableGrads: batch shape changed ({n} id slots, staged for {self.cap})")
        stream = _lib.stream_ptr(g.device)
        st = self.stage[self.step & 1]
        L = _lib.lib()
        _lib.check(L.grb_rows_scale(g.data_ptr(), self.touched.data_ptr(), n, D, V, 0, 1.0 / self.world, stream))
        if self.world > 1:
            _lib.check(L.grb_p2p_put_table_rows(
                g.data_ptr(), self.touched.data_ptr(), n, D, V, 0, 1.0, st["dst_rows"], st["dst_ids"],
                self.world - 1, self.rank * n, stream))

    def add_peers(self, param: torch.nn.Parameter) -> None:
        from . import _lib
        g = param.grad
        V, D = g.shape
        st = self.stage[self.step & 1]
        self.step += 1
        _lib.check(_lib.lib().grb_rows_scatter_add(
            st["rows"].data_ptr(), D, st["ids"].data_ptr(), g.data_ptr(), self.world * self.cap, D, V, 0,
            _lib.stream_ptr(g.device)))
        self.touched = None


class PeerGradients:
    """Data-parallel gradient reduction of a whole model over peer memory (no DistributedDataParallel,
    no NCCL call on the data path): the reference's `strategy: ddp` (configs/trainer/ddp.yaml:4) as two
    kernels of this package between two stream barriers per step.

    * item table (in-batch sampler): sparse rows, ``PeerTableGrads.put`` / ``add_peers``;
    * every other parameter: the gradients are gathered into one symmetric fp32 buffer and
      all-reduced in place by ``grb_p2p_allreduce`` (two-shot: each rank sums its 1 / world slice of
      every rank's buffer with peer loads and stores the average into every rank's buffer with peer
      stores); ``param.grad`` then views that buffer.

    ``reduce()`` goes between ``loss.backward()`` and the optimizer.  Under DDP the all-reduce of a
    graph-replayed backward cannot overlap anything (the whole backward is one autograd node, every
    bucket fires at its end) and costs 0.13 ms of launch + ring latency per step for 5 MB; this costs
    the barriers (the slowest rank's arrival) plus ~20 us.  Sums run in rank order, so all ranks hold
    bit-identical gradients."""

    def __init__(self, model: "RetrievalModel", group=None) -> None:
        import torch.distributed as dist
        from .peer import PeerBarrier, symmetric_empty
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        table = model.embeddings._item_emb.weight
        dev = table.device
        self.sync = PeerBarrier(group, dev)
        self.table = None
        if isinstance(model.negatives_sampler, InBatchNegativesSampler):
            self.table = PeerTableGrads(table, group, hook=False, sync=self.sync)
            object.__setattr__(model, "_peer_grads", self.table)
        self.dense = [p for p in model.parameters() if p.requires_grad and (self.table is None or p is not table)]
        src = dist.get_global_rank(self.group, 0)
        for p in self.dense:
            dist.broadcast(p.data, src=src, group=group)
        sizes = [-(-p.numel() // 4) * 4 for p in self.dense]        # every view 16-byte aligned
        total = sum(sizes)
        self.flat, self._h, self.ptrs = symmetric_empty((max(total, 4),), torch.float32, dev, self.group)
        self.flat.zero_()
        self.views, o = [], 0
        for p, sz in zip(self.dense, sizes):
            self.views.append(self.flat[o:o + p.numel()].view(p.shape))
            o += sz
        self.numel = total
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)

    @torch.no_grad()
    def reduce(self) -> None:
        from . import _lib
        dev = self.flat.device
        if self.table is not None:
            self.table.put(self.table.weight)
        have = [(v, p.grad) for v, p in zip(self.views, self.dense) if p.grad is not None]
        for v, p in zip(self.views, self.dense):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        self.sync.barrier(0, dev)          # peers' table rows have landed; every rank's flat buffer is filled
        if self.table is not None:
            self.table.add_peers(self.table.weight)
        _lib.check(_lib.lib().grb_p2p_allreduce(self.ptrs, self.world, self.rank, self.numel, 1.0 / self.world,
                                                _lib.stream_ptr(dev)))
        self.sync.barrier(1, dev)          # every slice has landed everywhere
        for v, p in zip(self.views, self.dense):
            p.grad = v


class _StepStack(torch.nn.Module):
    """What ``enable_step_graphs`` captures: device inputs -> loss, at a fixed row bucket.  Holds
    the model without registering it as a child."""

    def __init__(self, model: "RetrievalModel", t_pad: int) -> None:
        super().__init__()
        object.__setattr__(self, "_model", model)
        self._t_pad = t_pad

    def parameters(self, recurse: bool = True):
        return self._model.parameters(recurse)

    def forward(self, lengths, ids, ts, target_ids, target_ts):
        return self._model._loss_impl(lengths, ids, ts, target_ids, target_ts,
                                      total_length=self._t_pad, padded=True)


class RetrievalModel(torch.nn.Module):
    """The retrieval task without Lightning: ``training_loss`` == the body of
    Retrieval.training_step up to the loss (retrieval.py:80-133); ``retrieve`` == :21-48."""

    def __init__(self, cfg: RetrievalConfig, all_item_ids: torch.Tensor, sharded_index: bool = False):
        super().__init__()
        self.cfg = cfg
        D = cfg.embedding_dim
        self.embeddings = ItemEmbeddings(cfg.num_items, D, cfg.split_year_embedding)
        self.preprocessor = PositionalPreprocessor(cfg.N, D, cfg.dropout)
        self.sequence_encoder = HSTU(
            max_sequence_len=cfg.max_sequence_length, max_output_len=cfg.gr_output_length + 1,
            embedding_dim=D, item_embedding_dim=D, num_blocks=cfg.num_blocks,
            num_heads=cfg.num_heads, linear_dim=cfg.linear_dim, attention_dim=cfg.attention_dim,
            normalization="rel_bias", linear_config="uvqk", linear_activation="silu",
            linear_dropout_rate=cfg.dropout, attn_dropout_rate=0.0,
            compute_dtype=cfg.compute_dtype)
        self.postprocessor = L2NormPostprocessor(D, cfg.l2_eps)
        self.similarity = DotProductSimilarity()
        if cfg.sampler == "local":
            self.negatives_sampler = LocalNegativesSampler(
                l2_norm=True, l2_norm_eps=cfg.l2_eps, all_item_ids=all_item_ids.tolist())
        else:
            self.negatives_sampler = InBatchNegativesSampler(
                l2_norm=True, l2_norm_eps=cfg.l2_eps, dedup_embeddings=True)
            self.negatives_sampler.max_item_id = cfg.num_items     # enables the sort-free cache build
        self.loss = SampledSoftmaxLoss(cfg.num_negatives, cfg.temperature,
                                       bf16_backward=cfg.compute_dtype == torch.bfloat16)
        index_cls = ShardedCandidateIndex if sharded_index else CandidateIndex
        self.candidate_index = index_cls(k=cfg.top_k, ids=all_item_ids,
                                         top_k_module=MIPSBruteForceTopK())
        self._step_graph_rows = 0
        self._step_graph_lazy = True
        object.__setattr__(self, "_step_graphs", {})
        object.__setattr__(self, "_peer_grads", None)

    def enable_cuda_graphs(self, row_granularity: int = 1024, lazy: bool = True) -> None:
        """Training only: run the HSTU layer stack as captured CUDA graphs (see
        HSTUJagged.enable_cuda_graphs).  ``training_loss`` must then be given ``total_length``."""
        self.sequence_encoder._hstu.enable_cuda_graphs(row_granularity, lazy)

    def precapture_cuda_graphs(self, total_lengths, batch_size: int) -> int:
        """Capture the graphs for the row buckets of ``total_lengths`` now (required before the
        model is wrapped in DistributedDataParallel)."""
        enc = self.sequence_encoder
        dev = self.embeddings._item_emb.weight.device
        return enc._hstu.precapture(total_lengths, batch_size, self.cfg.N, self.cfg.embedding_dim, dev,
                                    enc._attn_mask, self.embeddings._item_emb.weight.dtype)

    def disable_cuda_graphs(self) -> None:
        self.sequence_encoder._hstu.disable_cuda_graphs()
        self._step_graph_rows = 0

    def enable_step_graphs(self, row_granularity: int = 1024, lazy: bool = True) -> None:
        """Training only: capture the WHOLE loss computation (embedding lookup, sampler cache,
        pre-processor, HSTU stack, post-processor, sampled-softmax loss) and its backward as one
        pair of CUDA graphs per padded row count.  Every jagged tensor is padded with zero rows to
        the bucket, the in-batch cache is padded and masked by device-side counts, so nothing in the
        step has a data-dependent shape.  The host side of a step is then two graph launches, the
        optimizer and the input copies.  ``training_loss`` must be given ``total_length``.
        Under DistributedDataParallel capture with ``precapture_step_graphs`` before wrapping."""
        hstu = self.sequence_encoder._hstu
        hstu.disable_cuda_graphs(drop=True)               # the stack is captured as part of the step
        hstu.enable_cuda_graphs(row_granularity, lazy=False)
        self._step_graph_rows = int(row_granularity)
        self._step_graph_lazy = bool(lazy)

    def enable_peer_table_grads(self, group=None):
        """Multi-GPU training: reduce the item-table gradient over peer memory (PeerTableGrads)
        instead of DDP's dense all-reduce.  Call after the model is on its device and BEFORE
        wrapping in DistributedDataParallel; returns the parameter names (relative to this module)
        that DDP must be told to ignore."""
        if not isinstance(self.negatives_sampler, InBatchNegativesSampler):
            # LocalNegativesSampler reads (and back-propagates into) random rows of the whole table:
            # the rows a step touches are then not the batch's ids and the sparse exchange is moot
            raise NotImplementedError("enable_peer_table_grads needs the in-batch negatives sampler")
        object.__setattr__(self, "_peer_grads", PeerTableGrads(self.embeddings._item_emb.weight, group))
        return ["embeddings._item_emb.weight"]

    def enable_peer_gradients(self, group=None) -> "PeerGradients":
        """Multi-GPU training without DistributedDataParallel: returns the reducer whose ``reduce()``
        goes between ``backward()`` and the optimizer step (PeerGradients)."""
        return PeerGradients(self, group)

    def precapture_step_graphs(self, rows, total_lengths) -> int:
        for row, tot in zip(rows, total_lengths):
            self._step_graph(self._device_inputs(row), int(tot), capture=True)
        return len(self._step_graphs)

    def _device_inputs(self, row):
        dev = self.embeddings._item_emb.weight.device
        return (row["history_lengths"].to(dev, non_blocking=True),
                row["historical_ids"].to(dev, non_blocking=True),
                row["historical_timestamps"].to(dev, non_blocking=True),
                row["target_ids"].to(dev, non_blocking=True).view(-1, 1),
                row["target_timestamps"].to(dev, non_blocking=True).view(-1, 1))

    def _step_graph(self, inputs, total_length: int, capture: bool):
        t_pad = -(-max(total_length, 1) // self._step_graph_rows) * self._step_graph_rows
        key = (t_pad,) + tuple((tuple(t.shape), t.dtype) for t in inputs)
        run = self._step_graphs.get(key)
        if run is None and capture:
            # a live autograd graph from an earlier eager step (the sampler's cache holds one) keeps
            # the parameters' gradient accumulators bound to the stream they were made on, and a
            # capture may not touch the default stream: drop it first
            for name in ("_cached_embeddings", "_cached_ids", "_cached_count"):
                if hasattr(self.negatives_sampler, name):
                    setattr(self.negatives_sampler, name, None)
            stack = _StepStack(self, t_pad)
            run = torch.cuda.make_graphed_callables(stack, tuple(t.clone() for t in inputs))
            self._step_graphs[key] = run
        return run, t_pad

    # generative_recommenders.py:355-393
    def forward(self, sf: SequentialFeatures, total_length: Optional[int] = None,
                jagged_output: bool = False) -> torch.Tensor:
        """(B, N, D) normalised encodings; ``jagged_output``: the (T, D) valid rows only (the
        post-processor acts per row, so it can run on the jagged rows directly)."""
        lengths, x, valid, _ = self.preprocessor(sf.past_lengths, sf.past_ids, sf.past_embeddings,
                                                 sf.past_payloads)
        x, _ = self.sequence_encoder(past_lengths=lengths, user_embeddings=x, valid_mask=valid,
                                     past_payloads=sf.past_payloads, total_length=total_length,
                                     jagged_output=jagged_output)
        if jagged_output and total_length is not None and x.size(0) != total_length:
            x = x[:total_length]       # drop the zero rows of the CUDA-graph row bucket
        return self.postprocessor(x)

    def training_loss(self, row: Dict[str, torch.Tensor], total_length: Optional[int] = None) -> torch.Tensor:
        inputs = self._device_inputs(row)
        if self._peer_grads is not None and self.training and torch.is_grad_enabled():
            self._peer_grads.note_ids(inputs[1], inputs[3])      # history ids, target ids
        if (self._step_graph_rows and total_length is not None and self.training
                and torch.is_grad_enabled() and inputs[0].is_cuda):
            run, t_pad = self._step_graph(inputs, int(total_length), capture=self._step_graph_lazy)
            if run is not None:
                return run(*inputs)
            return self._loss_impl(*inputs, total_length=t_pad, padded=True)
        return self._loss_impl(*inputs, total_length=total_length, padded=False)

    def _loss_impl_jagged(self, lengths, ids, ts, target_ids, target_ts, total_length: int,
                          padded: bool) -> Optional[torch.Tensor]:
        """The same step without any padded (B, N, D) tensor: every consumer of
        ``input_embeddings`` -- the pre-processor, the in-batch cache, the supervision embeddings -- reads
        the table rows of the VALID positions directly (the pre-processor fused with the gather, the
        positional embedding, dropout and the cast to the compute dtype in one kernel).  One fp32
        table, in-batch sampler with a small id space or the local sampler; None = not applicable."""
        c = self.cfg
        table = self.embeddings._item_emb.weight
        if (self.embeddings._year_emb is not None or not table.is_cuda or table.dtype != torch.float32
                or c.embedding_dim % 4 or total_length is None):
            return None
        sf, target_ids = _features_on_device(lengths, ids, ts, target_ids, target_ts, c.gr_output_length + 1)
        sf.past_ids.scatter_(dim=1, index=sf.past_lengths.view(-1, 1), src=target_ids.view(-1, 1))
        sup_ids = sf.past_ids
        off = ops.asynchronous_complete_cumsum(sf.past_lengths)
        tot = total_length
        # every reader of the table below scatters its gradient into ONE dense buffer (GF.TableGradScope)
        scope = GF.TableGradScope(table)
        if isinstance(self.negatives_sampler, InBatchNegativesSampler):
            # the valid ids of row b are its first length + 1 entries (history + target)
            if not self.negatives_sampler.process_batch_table(
                    sup_ids, off, tot + sup_ids.size(0), table, padded=padded, grad_scope=scope, rows_extra=1):
                return None
        else:
            self.negatives_sampler._embeddings_module = self.embeddings
        p_drop = c.dropout if self.training else 0.0
        seed = (torch.randint(0, 2 ** 62, (1,), device=table.device, dtype=torch.int64) if p_drop > 0 else None)
        # layer-stack CUDA graphs (enable_cuda_graphs): the encoder wants its rows padded to a bucket
        g_rows = self.sequence_encoder._hstu._graph_rows
        enc_rows = tot if (padded or not g_rows or not self.training) else -(-max(tot, 1) // g_rows) * g_rows
        xj = GF.jagged_input(table, self.preprocessor._pos_emb.weight, sup_ids, off, enc_rows,
                             c.embedding_dim ** 0.5, p_drop, seed,
                             out_dtype=c.compute_dtype or torch.float32, grad_scope=scope)
        enc, _ = self.sequence_encoder(past_lengths=sf.past_lengths, user_embeddings=xj, valid_mask=None,
                                       past_payloads=sf.past_payloads, total_length=tot,
                                       jagged_output=True, rows_padded=padded or enc_rows != tot)
        if enc.size(0) != tot:
            enc = enc[:tot]            # drop the zero rows of the encoder's row bucket
        out_rows = self.postprocessor(enc)
        sup_ids_j = ops.dense_to_jagged(sup_ids[:, 1:], off, total=tot, zero_tail=padded)
        jag = dict(
            output_embeddings=out_rows,
            supervision_ids=sup_ids_j,
            supervision_embeddings=GF.embedding_lookup(table, sup_ids_j, 0, grad_scope=scope),
            supervision_weights=(sup_ids_j != 0).float(),
        )
        return self.loss.jagged_forward(negatives_sampler=self.negatives_sampler,
                                        similarity=self.similarity, **jag)

    def _loss_impl(self, lengths, ids, ts, target_ids, target_ts, total_length: Optional[int],
                   padded: bool) -> torch.Tensor:
        """``padded``: ``total_length`` is a row bucket >= sum(lengths); every jagged tensor has that
        many rows, the ones past the real total being zero (and weighing zero in the loss)."""
        import os
        if os.environ.get("GRB_NO_FUSED_INPUT") != "1":
            fused = self._loss_impl_jagged(lengths, ids, ts, target_ids, target_ts, total_length, padded)
            if fused is not None:
                return fused
        sf, target_ids = _features_on_device(lengths, ids, ts, target_ids, target_ts,
                                             self.cfg.gr_output_length + 1)
        sf.past_ids.scatter_(dim=1, index=sf.past_lengths.view(-1, 1), src=target_ids.view(-1, 1))
        input_emb = self.embeddings.get_item_embeddings(sf.past_ids)
        sf = sf._replace(past_embeddings=input_emb)
        sup_ids = sf.past_ids
        off = ops.asynchronous_complete_cumsum(sf.past_lengths)
        tot = total_length
        # The sampler's cache is built before the encoder runs (the reference does it after,
        # retrieval.py:117-123; neither consumes RNG nor depends on the other) and with static
        # shapes: the step then contains no device synchronisation at all, so the host can
        # enqueue step k+1 while the GPU still runs step k.
        if isinstance(self.negatives_sampler, InBatchNegativesSampler):
            if tot is not None:
                # valid ids are the first length+1 entries of each row; the embeddings the
                # reference looks up again (get_item_embeddings(flat)) are input_emb itself
                self.negatives_sampler.process_batch_prefix(
                    sup_ids, input_emb, off + torch.arange(off.numel(), device=off.device, dtype=off.dtype),
                    tot + sup_ids.size(0), static_shapes=True, padded=padded)
            else:
                flat = sup_ids.view(-1)
                self.negatives_sampler.process_batch(
                    ids=flat, presences=(flat != 0), embeddings=input_emb.view(flat.numel(), -1))
        else:
            self.negatives_sampler._embeddings_module = self.embeddings
        # generative_recommenders.py:407-425 (ids go through float32 there; exact below 2^24 —
        # here they are gathered as integers, which is exact everywhere).  The reference pads the
        # encoder output to (B, N, D), normalises it and packs rows [0, length) again; those are
        # exactly the encoder's jagged rows, so they are normalised and used as they are.
        out_rows = self.forward(sf, total_length, jagged_output=True)
        jag = dict(
            output_embeddings=out_rows,
            supervision_ids=ops.dense_to_jagged(sup_ids[:, 1:], off, total=tot, zero_tail=padded),
            supervision_embeddings=ops.dense_to_jagged(input_emb[:, 1:, :], off, total=tot, zero_tail=padded),
            supervision_weights=ops.dense_to_jagged((sup_ids[:, 1:] != 0).float(), off, total=tot,
                                                    zero_tail=padded),
        )
        return self.loss.jagged_forward(negatives_sampler=self.negatives_sampler,
                                        similarity=self.similarity, **jag)

    @torch.no_grad()
    def refresh_index(self) -> None:
        """retrieval.py:165-169: re-embed + normalise the whole corpus."""
        emb = self.negatives_sampler.normalize_embeddings(
            self.embeddings.get_item_embeddings(self.candidate_index.ids))
        self.candidate_index.update_embeddings(emb)

    @torch.inference_mode()
    def retrieve(self, row: Dict[str, torch.Tensor], filter_past_ids: bool = True):
        dev = self.embeddings._item_emb.weight.device
        sf, _ = seq_features_from_row(row, dev, self.cfg.gr_output_length + 1)
        sf = sf._replace(past_embeddings=self.embeddings.get_item_embeddings(sf.past_ids))
        cur = ops.get_current_embeddings(sf.past_lengths, self.forward(sf))
        if self.candidate_index.embeddings is None:
            self.refresh_index()
        return self.candidate_index.get_top_k_outputs(
            query_embeddings=cur, invalid_ids=(sf.past_ids if filter_past_ids else None))


def synthetic_batch(cfg: RetrievalConfig, all_item_ids: torch.Tensor, batch_size: int, seed: int,
                    min_len: int = 20, full: bool = False) -> Dict[str, torch.Tensor]:
    """SURVEY §8(d) generator: the dict ``seq_features_from_row`` consumes (CPU tensors)."""
    gen = torch.Generator().manual_seed(seed)
    L = cfg.max_sequence_length
    lengths = torch.full((batch_size,), L) if full else \
        torch.randint(min_len, L + 1, (batch_size,), generator=gen)
    pick = torch.randint(0, all_item_ids.numel(), (batch_size, L), generator=gen)
    ids = all_item_ids[pick]
    ts = 978_300_000 + torch.cumsum(torch.randint(1, 5000, (batch_size, L), generator=gen), dim=1)
    valid = torch.arange(L).unsqueeze(0) < lengths.unsqueeze(1)
    last_ts = torch.gather(ts, 1, (lengths - 1).clamp(min=0).view(-1, 1)).squeeze(1)
    tgt = all_item_ids[torch.randint(0, all_item_ids.numel(), (batch_size,), generator=gen)]
    return {
        "history_lengths": lengths,
        "historical_ids": ids * valid,
        "historical_timestamps": ts * valid,
        "target_ids": tgt,
        "target_timestamps": last_ts + 100,
    }


def synthetic_item_ids(num_distinct: int, max_id: int, seed: int = 42) -> torch.Tensor:
    gen = torch.Generator().manual_seed(seed)
    return (torch.randperm(max_id, generator=gen)[:num_distinct] + 1).sort().values
