"""HSTU encoder — drop-in for the reference's ``models/sequential_encoders/hstu.py``.

Same class names, constructor kwargs, forward signatures and parameter names
(``_hstu._attention_layers.{i}._uvqk``, ``._o.weight/.bias``, ``._rel_attn_bias._ts_w/._pos_w``,
buffer ``_attn_mask``) as /root/reference/src/generative_recommenders_pl/models/
sequential_encoders/hstu.py (:71 bias, :208 STU layer, :426 HSTUJagged, :521 HSTU), so Hydra
``_target_`` strings and checkpoints carry over.  The difference is below the module boundary:
the bias (:96-128), the attention (:134-205), both LayerNorms and the U-gate (:258-264, :300,
:402) run as fused jagged sm_100a kernels, so no padded q/k/v, no (B,H,N,N) scores and no
(B,N,N) bucket tensor ever exist.

Extension: ``HSTU(..., compute_dtype=torch.bfloat16)`` runs the layers with bf16 activations
(fp32 master weights, fp32 accumulation; tcgen05 attention) — the reference is fp32-only
(autocast disabled, hstu.py:592).
"""
from __future__ import annotations

import abc
import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import functional as GF
from . import ops

TIMESTAMPS_KEY = "timestamps"

HSTUCacheState = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]


def tabulate_bucket_thresholds(bucketization_fn: Callable[[torch.Tensor], torch.Tensor],
                               num_buckets: int) -> torch.Tensor:
    """thresholds[t] = smallest d >= 0 with clamp(fn(d), 0, num_buckets) >= t+1  (int64, CPU).

    The reference evaluates ``bucketization_fn`` on every (i, j) pair each layer
    (hstu.py:117-123).  For a monotone fn of |d| the bucket is ``#{t: thresholds[t] <= |d|}``;
    the table is built by bisection on the *same* fn evaluated by torch on the CPU, so the
    kernel's integer compare reproduces the reference's float32 log/div/trunc bit for bit.
    """
    hi_cap = (1 << 62)
    probe = torch.tensor([0, 1, 2, 3, 7, 100, 12345, 10 ** 9, 10 ** 12], dtype=torch.int64)
    pos = torch.clamp(bucketization_fn(probe), 0, num_buckets)
    neg = torch.clamp(bucketization_fn(-probe), 0, num_buckets)
    if not torch.equal(pos, neg):
        raise NotImplementedError("bucketization_fn must depend on |x| only")
    if (pos[1:] < pos[:-1]).any():
        raise NotImplementedError("bucketization_fn must be non-decreasing in |x|")
    targets = torch.arange(1, num_buckets + 1, dtype=torch.int64)
    lo = torch.zeros(num_buckets, dtype=torch.int64)            # fn(lo) < target (or lo == 0)
    hi = torch.full((num_buckets,), hi_cap, dtype=torch.int64)  # fn(hi) >= target, or unreachable
    reach = torch.clamp(bucketization_fn(hi), 0, num_buckets) >= targets
    zero_ok = torch.clamp(bucketization_fn(lo), 0, num_buckets) >= targets
    for _ in range(64):
        mid = lo + (hi - lo) // 2
        ge = torch.clamp(bucketization_fn(mid), 0, num_buckets) >= targets
        hi = torch.where(ge, mid, hi)
        lo = torch.where(ge, lo, mid)
    thr = torch.where(zero_ok, torch.zeros_like(hi), hi)
    thr = torch.where(reach, thr, torch.full_like(thr, torch.iinfo(torch.int64).max))
    return thr


class RelativeAttentionBiasModule(torch.nn.Module):
    @abc.abstractmethod
    def forward(self, all_timestamps: torch.Tensor) -> torch.Tensor:
        """all_timestamps: (B, N) int64 -> float tensor broadcastable to (B, N, N)."""


class RelativePositionalBias(RelativeAttentionBiasModule):
    """bias[i, j] = w[N-1+j-i]  (reference hstu.py:50-68)."""

    def __init__(self, max_seq_len: int) -> None:
        super().__init__()
        self._max_seq_len: int = max_seq_len
        self._w = torch.nn.Parameter(torch.empty(2 * max_seq_len - 1).normal_(mean=0, std=0.02))

    def forward(self, all_timestamps: torch.Tensor) -> torch.Tensor:
        n = self._max_seq_len
        ar = torch.arange(n, device=self._w.device)
        return self._w[(n - 1) + ar.view(1, n) - ar.view(n, 1)].unsqueeze(0)


class RelativeBucketedTimeAndPositionBasedBias(RelativeAttentionBiasModule):
    """pos_w[N-1+j-i] + ts_w[bucket(ts[i+1] - ts[j])]  (reference hstu.py:71-128).

    Inside the encoder this module is only a parameter holder: the fused attention kernel
    reads ``_ts_w``, ``_pos_w`` and the tabulated thresholds directly.  ``forward`` still
    returns the dense (B, N, N) bias for callers that want it."""

    def __init__(self, max_seq_len: int, num_buckets: int,
                 bucketization_fn: Callable[[torch.Tensor], torch.Tensor]) -> None:
        super().__init__()
        self._max_seq_len: int = max_seq_len
        self._ts_w = torch.nn.Parameter(torch.empty(num_buckets + 1).normal_(mean=0, std=0.02))
        self._pos_w = torch.nn.Parameter(torch.empty(2 * max_seq_len - 1).normal_(mean=0, std=0.02))
        self._num_buckets: int = num_buckets
        self._bucketization_fn = bucketization_fn
        self.register_buffer("_bucket_thresholds",
                             tabulate_bucket_thresholds(bucketization_fn, num_buckets),
                             persistent=False)

    def forward(self, all_timestamps: torch.Tensor) -> torch.Tensor:
        N = self._max_seq_len
        ar = torch.arange(N, device=all_timestamps.device)
        pos = self._pos_w[(N - 1) + ar.view(1, N) - ar.view(N, 1)]
        nxt = torch.cat([all_timestamps[:, 1:], all_timestamps[:, N - 1:N]], dim=1)
        d = (nxt.unsqueeze(2) - all_timestamps.unsqueeze(1)).abs()
        buckets = torch.bucketize(d, self._bucket_thresholds, right=True)
        return pos.unsqueeze(0) + self._ts_w[buckets]


class SequentialTransductionUnitJagged(torch.nn.Module):
    """One HSTU layer on jagged rows (reference hstu.py:208-423)."""

    def __init__(self, embedding_dim: int, linear_hidden_dim: int, attention_dim: int,
                 dropout_ratio: float, attn_dropout_ratio: float, num_heads: int,
                 linear_activation: str,
                 relative_attention_bias_module: Optional[RelativeAttentionBiasModule] = None,
                 normalization: str = "rel_bias", linear_config: str = "uvqk",
                 concat_ua: bool = False, epsilon: float = 1e-6,
                 max_length: Optional[int] = None) -> None:
        super().__init__()
        self._embedding_dim: int = embedding_dim
        self._linear_dim: int = linear_hidden_dim
        self._attention_dim: int = attention_dim
        self._dropout_ratio: float = dropout_ratio
        self._attn_dropout_ratio: float = attn_dropout_ratio  # stored, never applied (as :230)
        self._num_heads: int = num_heads
        self._rel_attn_bias: Optional[RelativeAttentionBiasModule] = relative_attention_bias_module
        self._normalization: str = normalization
        self._linear_config: str = linear_config
        if linear_config != "uvqk":
            raise ValueError(f"Unknown linear_config {linear_config}")
        self._uvqk = torch.nn.Parameter(
            torch.empty((embedding_dim,
                         linear_hidden_dim * 2 * num_heads + attention_dim * num_heads * 2)
                        ).normal_(mean=0, std=0.02))
        self._linear_activation: str = linear_activation
        self._concat_ua: bool = concat_ua
        self._o = torch.nn.Linear(
            in_features=linear_hidden_dim * num_heads * (3 if concat_ua else 1),
            out_features=embedding_dim)
        torch.nn.init.xavier_uniform_(self._o.weight)
        self._eps: float = epsilon

    def _norm_input(self, x: torch.Tensor) -> torch.Tensor:
        return GF.layer_norm_gate(x, None, self._eps)

    def _softmax_branch(self, x, u, v, q, k, x_offsets, all_timestamps, invalid_attn_mask,
                        return_cache_states):
        """normalization="softmax_rel_bias" (hstu.py:337-384): ONE softmax attention over the full
        H*dqk width (no head split), bias added before the softmax, causal mask applied AFTER it, so a
        row's weights do not sum to one.  Not a hot path of any shipped config: composed from the jagged
        kernels, cuBLAS bmm and ATen softmax in the reference's padded formulation, O(B N^2) memory."""
        n = invalid_attn_mask.size(-1)
        padded_q = ops.jagged_to_padded_dense(q.contiguous(), x_offsets, n, 0.0)
        padded_k = ops.jagged_to_padded_dense(k.contiguous(), x_offsets, n, 0.0)
        qk = torch.bmm(padded_q, padded_k.transpose(1, 2))
        if self._rel_attn_bias is not None:
            qk = qk + self._rel_attn_bias(all_timestamps).to(qk.dtype)
        # as on the fused path only the size of invalid_attn_mask is read: the mask is the causal
        # lower triangle HSTU registers (hstu.py:595-607, :667)
        causal = torch.ones(n, n, dtype=qk.dtype, device=qk.device).tril_()
        qk = F.softmax(qk / math.sqrt(self._attention_dim), dim=-1) * causal
        attn_output = ops.dense_to_jagged(
            torch.bmm(qk, ops.jagged_to_padded_dense(v.contiguous(), x_offsets, n, 0.0)), x_offsets)
        if self._concat_ua:
            a = self._norm_attn_output(attn_output)
            o_input = torch.cat([u, a, u * a], dim=-1)
        else:
            o_input = GF.layer_norm_gate(attn_output, u.contiguous(), self._eps)
        new_outputs = GF.master_linear(
            F.dropout(o_input, p=self._dropout_ratio, training=self.training),
            self._o.weight, self._o.bias) + x
        cache_state = (v.contiguous(), padded_q, padded_k, new_outputs) if return_cache_states else None
        return new_outputs, cache_state

    def _norm_attn_output(self, x: torch.Tensor) -> torch.Tensor:
        return GF.layer_norm_gate(x, None, self._eps)

    def forward(self, x: torch.Tensor, x_offsets: torch.Tensor,
                all_timestamps: Optional[torch.Tensor], invalid_attn_mask: torch.Tensor,
                delta_x_offsets: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                cache: Optional[HSTUCacheState] = None,
                return_cache_states: bool = False,
                bucket_cache: Optional[torch.Tensor] = None,
                rows_padded: bool = False,
                dropout_seed: Optional[torch.Tensor] = None,
                aux: Optional[dict] = None):
        """x: (sum_i N_i, D); x_offsets: (B+1); all_timestamps: (B, N) int64 or None;
        invalid_attn_mask: (N, N) — only its size is read: the kernel applies the causal
        lower-triangular mask that HSTU registers (hstu.py:595-607,667)."""
        # Incremental path (hstu.py:293-298): x, u, v, q, k are restricted to the rows
        # delta_x_offsets[0] (one new position per sequence); the caches of the earlier positions
        # come from a previous call with return_cache_states=True.  A cache without
        # delta_x_offsets is ignored, as in the reference.
        incremental = delta_x_offsets is not None
        if incremental:
            if cache is None:
                raise ValueError("delta_x_offsets needs this layer's cache state (hstu.py:295)")
            rows = delta_x_offsets[0].long()
            x = x[rows, :]
            cached_v, cached_q, cached_k, cached_outputs = cache
        if self._normalization not in ("rel_bias", "hstu_rel_bias", "softmax_rel_bias"):
            raise ValueError(f"Unknown normalization method {self._normalization}")
        softmax = self._normalization == "softmax_rel_bias"
        if softmax and incremental:
            # the reference's own branch cannot run: `B = x_offsets.size() - 1` (hstu.py:339) is a
            # torch.Size minus an int
            raise NotImplementedError("softmax_rel_bias has no incremental path (nor does the reference: "
                                      "hstu.py:339 raises TypeError)")
        n: int = invalid_attn_mask.size(-1)
        H, dv, dqk = self._num_heads, self._linear_dim, self._attention_dim

        # (F.layer_norm(x), x): the residual operand of :413 leaves the same autograd node as the norm, so
        # that x receives ONE gradient (norm backward + residual, summed in the kernel)
        normed_x, x = GF.layer_norm_skip(x, self._eps)
        # training-mode dropout in front of the output projection (:404-408) is drawn inside the
        # u * norm(a) kernel from a per-step device seed; HSTUJagged hands one seed to all its layers
        drop = None
        if self.training and self._dropout_ratio > 0.0:
            if dropout_seed is None:
                dropout_seed = torch.randint(0, 2 ** 62, (1,), device=x.device, dtype=torch.int64)
            drop = (dropout_seed, getattr(self, "_layer_index", 0))
        sizes = [dv * H, dv * H, dqk * H, dqk * H]
        if self._linear_activation == "silu":
            # one tcgen05 GEMM with the SiLU in its epilogue (bf16 rows); cuBLAS + silu kernel otherwise
            # aux (from HSTUJagged): this layer's weights already in the compute dtype and zero-filled
            # accumulators for its weight gradients, each made by ONE launch for all layers
            aux = aux or {}
            u, v, q, k = GF.uvqk_projection(normed_x, self._uvqk, sizes, aux.get("w_uvqk"), aux.get("dw_uvqk"))
        elif self._linear_activation == "none":
            mm = GF.master_linear(normed_x, self._uvqk, None, w_in_out=True)
            u, v, q, k = torch.split(mm, sizes, dim=1)
        else:
            raise ValueError(f"Unknown linear_activation {self._linear_activation}")

        if softmax:
            return self._softmax_branch(x, u, v, q, k, x_offsets, all_timestamps, invalid_attn_mask,
                                        return_cache_states)
        bias = self._rel_attn_bias if all_timestamps is not None else None
        if not incremental and (
                (bias is not None and not isinstance(bias, RelativeBucketedTimeAndPositionBasedBias))
                or max(dqk, dv) > (128 if torch.is_grad_enabled() else 256)):
            # shapes / bias modules the fused kernels do not cover: heads wider than 128 (the reference's
            # default injection attention_dim = linear_dim = item_embedding_dim gives H = 1, d = D,
            # generative_recommenders.py:157-160) when a gradient is needed, wider than 256 at all,
            # or RelativePositionalBias (hstu.py:50-68).  The reference's padded formulation on the
            # jagged kernels + cuBLAS + ATen, O(B H N^2) memory: functional, not a hot path.
            attn_output = self._composite_attention(q, k, v, x_offsets, all_timestamps, n)
            return self._finish_layer(x, u, v, q, k, attn_output, x_offsets, n, False, None, None,
                                      return_cache_states, drop, aux)
        if bias is not None and not isinstance(bias, RelativeBucketedTimeAndPositionBasedBias):
            raise NotImplementedError(
                "the incremental path supports RelativeBucketedTimeAndPositionBasedBias only")
        bias_args = (all_timestamps if bias is not None else None,
                     bias._ts_w if bias is not None else None,
                     bias._pos_w if bias is not None else None,
                     bias._bucket_thresholds if bias is not None else None)
        if incremental:
            # hstu.py:321-322 and :151-177: the new rows go into the caches in place ...
            B = x_offsets.numel() - 1
            v = cached_v.index_copy_(0, rows, v)
            flat = delta_x_offsets[1].long() + torch.arange(0, B * n, n, device=x.device)
            padded_q = cached_q.view(B * n, -1).index_copy_(0, flat, q).view(B, n, -1)
            padded_k = cached_k.view(B * n, -1).index_copy_(0, flat, k).view(B, n, -1)
            # ... and where the reference recomputes the whole (B, H, N, N) attention to keep the
            # rows delta_x_offsets[0] (:179-204, :397-401), one row per sequence is computed
            attn_output = GF.hstu_attention_decode(
                q, padded_k, v, x_offsets, delta_x_offsets[1], *bias_args,
                N=n, num_heads=H, attention_dim=dqk, linear_dim=dv)
        else:
            attn_output = GF.hstu_attention(
                q, k, v, x_offsets, *bias_args,
                N=n, num_heads=H, attention_dim=dqk, linear_dim=dv, bucket_cache=bucket_cache,
                rows_padded=rows_padded)

        return self._finish_layer(x, u, v, q, k, attn_output, x_offsets, n, incremental,
                                  rows if incremental else None,
                                  (cached_outputs, padded_q, padded_k) if incremental else None,
                                  return_cache_states, drop, aux)

    def _composite_attention(self, q, k, v, x_offsets, all_timestamps, n: int) -> torch.Tensor:
        """hstu.py:179-204 as written: pad, (B, H, N, N) scores, + bias, SiLU / N, causal mask, P V, un-pad."""
        B = x_offsets.numel() - 1
        H, dv, dqk = self._num_heads, self._linear_dim, self._attention_dim
        pq = ops.jagged_to_padded_dense(q.contiguous(), x_offsets, n, 0.0).view(B, n, H, dqk)
        pk = ops.jagged_to_padded_dense(k.contiguous(), x_offsets, n, 0.0).view(B, n, H, dqk)
        pv = ops.jagged_to_padded_dense(v.contiguous(), x_offsets, n, 0.0).view(B, n, H, dv)
        qk = torch.einsum("bnhd,bmhd->bhnm", pq, pk)
        if all_timestamps is not None and self._rel_attn_bias is not None:
            qk = qk + self._rel_attn_bias(all_timestamps).unsqueeze(1).to(qk.dtype)
        qk = F.silu(qk) / n
        qk = qk * torch.ones(n, n, dtype=qk.dtype, device=qk.device).tril_()
        out = ops.dense_to_jagged(torch.einsum("bhnm,bmhd->bnhd", qk, pv).reshape(B, n, H * dv), x_offsets)
        if out.size(0) < q.size(0):        # fixed row buckets: the rows past offsets[-1] stay zero
            out = F.pad(out, (0, 0, 0, q.size(0) - out.size(0)))
        return out

    def _finish_layer(self, x, u, v, q, k, attn_output, x_offsets, n, incremental, rows, cached,
                      return_cache_states, drop=None, aux=None):
        if incremental:
            cached_outputs, padded_q, padded_k = cached
        if self._concat_ua:
            a = self._norm_attn_output(attn_output)
            o_input = torch.cat([u, a, u * a], dim=-1)
            o_input = F.dropout(o_input, p=self._dropout_ratio, training=self.training)
        elif drop is not None:
            o_input = GF.layer_norm_gate(attn_output, u, self._eps, self._dropout_ratio, *drop)
        else:
            o_input = GF.layer_norm_gate(attn_output, u, self._eps)

        aux = aux or {}
        new_outputs = GF.output_projection(o_input, self._o.weight, self._o.bias, x, aux.get("w_o"),
                                           aux.get("dw_o"), aux.get("db_o"))

        cache_state = None
        if incremental:
            new_outputs = cached_outputs.index_copy_(0, rows, new_outputs)   # hstu.py:415-418
            cache_state = (v, padded_q, padded_k, new_outputs)
        elif return_cache_states:
            cache_state = (
                v.contiguous(),
                ops.jagged_to_padded_dense(q.contiguous(), x_offsets, n, 0.0),
                ops.jagged_to_padded_dense(k.contiguous(), x_offsets, n, 0.0),
                new_outputs,
            )
        return new_outputs, cache_state


def _require_causal(invalid_attn_mask: torch.Tensor) -> None:
    """The kernels apply the causal lower triangle HSTU registers as ``_attn_mask`` (hstu.py:595-607, :667)
    themselves and read only the SIZE of ``invalid_attn_mask``; the reference multiplies by the tensor it is
    given (hstu.py:194).  Any other mask (full, windowed, target-aware) must not be silently ignored: it is
    checked once per mask tensor (one host comparison, remembered on the tensor object) and refused."""
    seen = getattr(invalid_attn_mask, "_grb_causal", None)
    if seen is not None and seen == invalid_attn_mask._version:
        return
    n = invalid_attn_mask.size(-1)
    m = invalid_attn_mask.reshape(-1, n)[-n:]
    # two spellings of the same mask: the reference's float `1 - _attn_mask` (non-zero = attend), or this
    # package's HSTU handing over its boolean `_attn_mask` buffer itself (True = masked, hstu.py:595-607)
    m = ~m if m.dtype == torch.bool else m != 0
    if not torch.equal(m, torch.ones(n, n, dtype=torch.bool, device=m.device).tril_()):
        raise NotImplementedError(
            "HSTU on the fused kernels supports the causal lower-triangular invalid_attn_mask only; "
            "got a different mask (the reference would multiply the scores by it, hstu.py:194)")
    invalid_attn_mask._grb_causal = invalid_attn_mask._version


class HSTUJagged(torch.nn.Module):
    """Stack of STU layers over jagged rows (reference hstu.py:426-518)."""

    def __init__(self, modules: List[SequentialTransductionUnitJagged],
                 autocast_dtype: Optional[torch.dtype]) -> None:
        super().__init__()
        self._attention_layers = torch.nn.ModuleList(modules=modules)
        self._autocast_dtype = autocast_dtype
        self._graph_rows = 0
        self._graph_lazy = True
        object.__setattr__(self, "_graphs", {})

    def _layer_aux(self, x: torch.Tensor) -> Optional[List[dict]]:
        """Per-layer operands every layer used to make for itself, one launch each: the bf16 copies of the
        fp32 master weights (uvqk, _o.weight) -> ONE multi-tensor cast; the zero-filled fp32 accumulators of
        the split-K weight / bias gradients -> ONE zero fill (only when a backward can follow)."""
        layers = list(self._attention_layers)
        if not (layers and x.is_cuda and x.dtype == torch.bfloat16):
            return None
        if not all(l._linear_activation == "silu" and l._uvqk.dtype == torch.float32 and l._uvqk.is_contiguous()
                   and l._o.weight.dtype == torch.float32 and l._o.weight.is_contiguous() for l in layers):
            return None
        L = len(layers)
        with torch.no_grad():
            sh = GF.cast_many_bf16([l._uvqk for l in layers] + [l._o.weight for l in layers])
        aux = [{"w_uvqk": sh[i], "w_o": sh[L + i]} for i in range(L)]
        if torch.is_grad_enabled() and any(l._uvqk.requires_grad or l._o.weight.requires_grad for l in layers):
            shapes = []
            for l in layers:
                shapes += [tuple(l._uvqk.shape), tuple(l._o.weight.shape), (l._o.weight.shape[0],)]
            with torch.no_grad():
                z = GF.zeros_many(shapes, x.device)
            for i in range(L):
                aux[i].update(dw_uvqk=z[3 * i], dw_o=z[3 * i + 1])
                if layers[i]._o.bias is not None and layers[i]._o.bias.dtype == torch.float32:
                    aux[i]["db_o"] = z[3 * i + 2]
        return aux

    def jagged_forward(self, x: torch.Tensor, x_offsets: torch.Tensor,
                       all_timestamps: Optional[torch.Tensor], invalid_attn_mask: torch.Tensor,
                       delta_x_offsets=None, cache=None, return_cache_states: bool = False,
                       rows_padded: bool = False):
        cache_states: List[HSTUCacheState] = []
        _require_causal(invalid_attn_mask)
        in_dtype = x.dtype
        if self._autocast_dtype is not None and x.dtype != self._autocast_dtype:
            x = x.to(self._autocast_dtype)
        # the time buckets are the same for every layer, head and direction: tabulate them once
        # (tcgen05 path only: bf16 activations, 64-wide heads)
        bucket_cache = None
        first = self._attention_layers[0] if len(self._attention_layers) else None
        if (first is not None and x.dtype == torch.bfloat16 and delta_x_offsets is None
                and first._attention_dim == 64 and first._linear_dim == 64
                and first._normalization != "softmax_rel_bias"):
            n_pad = invalid_attn_mask.size(-1)
            has_bias = (all_timestamps is not None
                        and isinstance(first._rel_attn_bias, RelativeBucketedTimeAndPositionBasedBias))
            # short sequences (N <= 256): masked tiles + item schedule for the persistent kernels,
            # needed with or without a relative bias
            short = GF.short_path_applies(x, 64, 64, n_pad)
            if has_bias or short:
                bucket_cache = GF.hstu_bucket_cache(
                    x_offsets, all_timestamps if has_bias else None,
                    first._rel_attn_bias._bucket_thresholds if has_bias else None, n_pad, masked=short)
        dropout_seed = None
        if self.training and any(l._dropout_ratio > 0.0 for l in self._attention_layers):
            dropout_seed = torch.randint(0, 2 ** 62, (1,), device=x.device, dtype=torch.int64)
        layer_aux = self._layer_aux(x) if delta_x_offsets is None else None
        for i, layer in enumerate(self._attention_layers):
            layer._layer_index = i
            x, cs = layer(x=x, x_offsets=x_offsets, all_timestamps=all_timestamps,
                          invalid_attn_mask=invalid_attn_mask, delta_x_offsets=delta_x_offsets,
                          cache=cache[i] if cache is not None else None,
                          return_cache_states=return_cache_states, bucket_cache=bucket_cache,
                          rows_padded=rows_padded, dropout_seed=dropout_seed,
                          aux=layer_aux[i] if layer_aux is not None else None)
            if return_cache_states:
                cache_states.append(cs)
        if x.dtype != in_dtype:
            x = x.to(in_dtype)
        return x, cache_states

    def forward(self, x: torch.Tensor, x_offsets: torch.Tensor,
                all_timestamps: Optional[torch.Tensor], invalid_attn_mask: torch.Tensor,
                delta_x_offsets=None, cache=None, return_cache_states: bool = False,
                total_length: Optional[int] = None, jagged_output: bool = False,
                rows_padded: bool = False):
        """x: (B, N, D) padded or (T, D) jagged.  Returns (B, N, D), cache states; with
        ``jagged_output`` the (>= T, D) jagged rows instead (rows past offsets[-1] are padding).
        ``rows_padded``: a jagged x carries zero rows past offsets[-1] (fixed row buckets)."""
        n = invalid_attn_mask.size(1)
        graphable = (self._graph_rows and x.is_cuda and torch.is_grad_enabled() and x.requires_grad
                     and self.training and all_timestamps is not None and delta_x_offsets is None
                     and cache is None and not return_cache_states)
        if graphable and ((x.dim() == 3 and total_length is not None)
                          or (x.dim() == 2 and rows_padded and x.shape[0] % self._graph_rows == 0)):
            # CUDA-graph path: jagged rows padded with zero rows up to a fixed bucket, so that one
            # captured (forward, backward) pair serves every batch of that bucket
            if x.dim() == 3:
                t_pad = -(-max(int(total_length), 1) // self._graph_rows) * self._graph_rows
                xj = ops.dense_to_jagged(x, x_offsets, total=t_pad, zero_tail=True)
            else:
                xj = x                   # already jagged and padded to a bucket by the caller
            run = self._graphed_stack(xj, x_offsets, all_timestamps, invalid_attn_mask,
                                      capture=self._graph_lazy)
            if run is not None:
                yj = run(xj, x_offsets, all_timestamps)
            else:   # no graph for this bucket and lazy capture is off: same padded rows, eagerly
                yj, _ = self.jagged_forward(x=xj, x_offsets=x_offsets, all_timestamps=all_timestamps,
                                            invalid_attn_mask=invalid_attn_mask, rows_padded=True)
            if jagged_output:
                return yj, []
            y = ops.jagged_to_padded_dense(values=yj, offsets=x_offsets, max_lengths=n,
                                           padding_value=0.0, padded_rows=True)
            return y, []
        if x.dim() == 3:
            x = ops.dense_to_jagged(x, x_offsets, total=total_length)
        jagged_x, cache_states = self.jagged_forward(
            x=x, x_offsets=x_offsets, all_timestamps=all_timestamps,
            invalid_attn_mask=invalid_attn_mask, delta_x_offsets=delta_x_offsets, cache=cache,
            return_cache_states=return_cache_states, rows_padded=rows_padded)
        if jagged_output:
            return jagged_x, cache_states
        y = ops.jagged_to_padded_dense(values=jagged_x, offsets=x_offsets,
                                       max_lengths=n, padding_value=0.0)
        return y, cache_states

    # ---- CUDA graphs (extension; the reference runs the layers eagerly, hstu.py:467-478) ----
    def enable_cuda_graphs(self, row_granularity: int = 1024, lazy: bool = True) -> None:
        """Run the layer stack (forward and backward) as captured CUDA graphs in training.  The
        ~40 launches per layer per direction of the eager path cost more host time than GPU time
        at the ml-20m shape; a graph replays them with one launch.  Jagged rows are padded with
        zero rows to a multiple of ``row_granularity`` and one graph pair is captured per padded
        size seen (lazily, on first use).  Needs ``total_length`` from the caller.

        Under DistributedDataParallel graphs must be captured BEFORE the model is wrapped (the
        wrapper pins the parameters' gradient accumulators to the stream it was built on, which a
        later capture may not touch): call ``precapture`` first and pass ``lazy=False`` so that
        sizes without a graph run eagerly."""
        self._graph_rows = int(row_granularity)   # graphs are keyed by shape: old ones stay valid
        self._graph_lazy = bool(lazy)

    def precapture(self, total_lengths, batch_size: int, n: int, width: int, device,
                   invalid_attn_mask: torch.Tensor, dtype: torch.dtype = torch.float32) -> int:
        """Captures the graph pair of every row bucket that ``total_lengths`` fall into, with
        synthetic inputs of the right shapes.  Returns the number of graphs held."""
        assert self._graph_rows > 0, "enable_cuda_graphs first"
        for t in sorted({-(-max(int(t), 1) // self._graph_rows) * self._graph_rows for t in total_lengths}):
            per = max(1, min(n - 1, t // batch_size))
            lengths = torch.full((batch_size,), per, dtype=torch.int64, device=device)
            offsets = ops.asynchronous_complete_cumsum(lengths)
            ts = torch.arange(n, device=device, dtype=torch.int64).repeat(batch_size, 1) * 1000 + 978_300_000
            xj = torch.zeros(t, width, device=device, dtype=dtype)
            self._graphed_stack(xj, offsets, ts, invalid_attn_mask, capture=True)
        return len(self._graphs)

    def disable_cuda_graphs(self, drop: bool = False) -> None:
        """Back to the eager path; captured graphs are kept for a later enable unless ``drop``."""
        self._graph_rows = 0
        if drop:
            object.__setattr__(self, "_graphs", {})

    def _graphed_stack(self, xj, x_offsets, timestamps, invalid_attn_mask, capture: bool = True):
        key = (tuple(xj.shape), xj.dtype, tuple(x_offsets.shape), x_offsets.dtype,
               tuple(timestamps.shape))
        run = self._graphs.get(key)
        if run is None and capture:
            stack = _LayerStack(self, invalid_attn_mask)
            sample = (xj.detach().clone().requires_grad_(True), x_offsets.clone(), timestamps.clone())
            run = torch.cuda.make_graphed_callables(stack, sample)
            self._graphs[key] = run
        return run


class _LayerStack(torch.nn.Module):
    """The callable that gets captured: all layers over padded jagged rows.  Holds the encoder
    without registering it as a child (it must not show up in the encoder's own state_dict)."""

    def __init__(self, encoder: "HSTUJagged", invalid_attn_mask: torch.Tensor) -> None:
        super().__init__()
        object.__setattr__(self, "_encoder", encoder)
        object.__setattr__(self, "_mask", invalid_attn_mask)

    def parameters(self, recurse: bool = True):
        return self._encoder.parameters(recurse)

    def forward(self, xj, x_offsets, timestamps):
        y, _ = self._encoder.jagged_forward(
            x=xj, x_offsets=x_offsets, all_timestamps=timestamps, invalid_attn_mask=self._mask,
            rows_padded=True)
        return y


def _default_bucketization(x: torch.Tensor) -> torch.Tensor:
    # the reference's lambda (hstu.py:579-581), kept as a named function so modules pickle
    return (torch.log(torch.abs(x).clamp(min=1)) / 0.301).long()


class HSTU(torch.nn.Module):
    """Top-level encoder the model config targets (reference hstu.py:521-672)."""

    def __init__(self, max_sequence_len: int, max_output_len: int, embedding_dim: int,
                 item_embedding_dim: int, num_blocks: int, num_heads: int, linear_dim: int,
                 attention_dim: int, normalization: str, linear_config: str,
                 linear_activation: str, linear_dropout_rate: float, attn_dropout_rate: float,
                 enable_relative_attention_bias: bool = True, concat_ua: bool = False,
                 compute_dtype: Optional[torch.dtype] = None) -> None:
        super().__init__()
        self._embedding_dim: int = embedding_dim
        self._item_embedding_dim: int = item_embedding_dim
        self._max_sequence_length: int = max_sequence_len
        self._num_blocks: int = num_blocks
        self._num_heads: int = num_heads
        self._dqk: int = attention_dim
        self._dv: int = linear_dim
        self._linear_activation: str = linear_activation
        self._linear_dropout_rate: float = linear_dropout_rate
        self._attn_dropout_rate: float = attn_dropout_rate
        self._enable_relative_attention_bias: bool = enable_relative_attention_bias
        n = max_sequence_len + max_output_len
        self._hstu = HSTUJagged(
            modules=[
                SequentialTransductionUnitJagged(
                    embedding_dim=embedding_dim, linear_hidden_dim=linear_dim,
                    attention_dim=attention_dim, normalization=normalization,
                    linear_config=linear_config, linear_activation=linear_activation,
                    num_heads=num_heads,
                    relative_attention_bias_module=(
                        RelativeBucketedTimeAndPositionBasedBias(
                            max_seq_len=n, num_buckets=128,
                            bucketization_fn=_default_bucketization)
                        if enable_relative_attention_bias else None),
                    dropout_ratio=linear_dropout_rate, attn_dropout_ratio=attn_dropout_rate,
                    concat_ua=concat_ua)
                for _ in range(num_blocks)
            ],
            autocast_dtype=compute_dtype,
        )
        self.register_buffer("_attn_mask",
                             torch.triu(torch.ones((n, n), dtype=torch.bool), diagonal=1))
        self.reset_params()

    def reset_params(self) -> None:
        # every encoder parameter name contains "_hstu", which the reference skips (:610-621)
        for name, params in self.named_parameters():
            if ("_hstu" in name) or ("_embedding_module" in name):
                continue
            try:
                torch.nn.init.xavier_normal_(params.data)
            except Exception:
                pass

    def debug_str(self) -> str:
        s = (f"HSTU-b{self._num_blocks}-h{self._num_heads}-dqk{self._dqk}-dv{self._dv}"
             f"-l{self._linear_activation}d{self._linear_dropout_rate}-ad{self._attn_dropout_rate}")
        if not self._enable_relative_attention_bias:
            s += "-norab"
        return s

    def forward(self, past_lengths: torch.Tensor, user_embeddings: torch.Tensor,
                valid_mask: torch.Tensor, past_payloads: Dict[str, torch.Tensor],
                delta_x_offsets=None, cache=None, return_cache_states: bool = False,
                total_length: Optional[int] = None, jagged_output: bool = False,
                rows_padded: bool = False):
        """past_lengths (B,) int; user_embeddings (B, N, D) (or jagged (T, D), extension); valid_mask
        unused (as :637); past_payloads["timestamps"] (B, N) int64.  Returns ((B, N, D), cache states);
        ``jagged_output`` (extension) skips the final jagged -> padded copy."""
        # only the mask's size is read downstream: pass the bool buffer itself instead of
        # materialising 1 - mask in the activation dtype every step (256 MB at N = 8192)
        return self._hstu(
            x=user_embeddings,
            x_offsets=ops.asynchronous_complete_cumsum(past_lengths),
            all_timestamps=past_payloads.get(TIMESTAMPS_KEY),
            invalid_attn_mask=self._attn_mask,
            delta_x_offsets=delta_x_offsets, cache=cache,
            return_cache_states=return_cache_states, total_length=total_length,
            jagged_output=jagged_output, rows_padded=rows_padded)
