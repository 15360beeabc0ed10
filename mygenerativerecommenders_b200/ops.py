"""Jagged operator layer — drop-in for the reference's ``models/utils/ops.py``.

Same nine function names, argument meaning and error behaviour as
``/root/reference/src/generative_recommenders_pl/models/utils/ops.py`` (:18, :41, :67, :117,
:149, :171, :190, :210, :229), but every function runs a hand-written sm_100a kernel from
``libgrb200.so`` through the C ABI (``include/grb200.h``).  CUDA tensors only — there is no
CPU / PyTorch fallback (the reference's fallbacks at ops.py:36-38, :60-64, :104-114 are what
this layer replaces).  ``install.py`` rebinds the reference module's attributes to these.

Extensions (keyword-only, optional): ``total`` on :func:`dense_to_jagged` lets a caller that
already knows ``offsets[-1]`` avoid the device->host read needed to size the output.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

__all__ = [
    "asynchronous_complete_cumsum", "dense_to_jagged", "jagged_to_padded_dense",
    "batch_gather_embeddings", "batch_scatter_embeddings", "get_current_embeddings",
    "jagged_or_dense_repeat_interleave_dim0", "jagged_or_dense_index_select_dim0",
    "mask_dense_by_aux_mask",
]


# --------------------------------------------------------------------------------------------
# raw kernel wrappers (no autograd)
# --------------------------------------------------------------------------------------------
def _cumsum_raw(lengths: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(lengths)
    bits = _lib.index_bits(lengths)
    lengths = lengths.contiguous()
    out = torch.empty(lengths.numel() + 1, dtype=lengths.dtype, device=lengths.device)
    _lib.check(_lib.lib().grb_complete_cumsum(
        lengths.data_ptr(), out.data_ptr(), lengths.numel(), bits, _lib.stream_ptr(lengths.device)))
    return out


def _dense_layout(dense: torch.Tensor) -> tuple[torch.Tensor, int, int]:
    """(tensor, row_bytes, batch_stride_bytes) with rows contiguous inside each batch entry."""
    B, N = dense.shape[0], dense.shape[1]
    row_elems = 1
    for s in dense.shape[2:]:
        row_elems *= s
    ok = dense.dim() >= 2
    # inner dims must be contiguous, dim-1 stride must equal the row size
    expect = 1
    for d in range(dense.dim() - 1, 1, -1):
        if dense.shape[d] != 1 and dense.stride(d) != expect:
            ok = False
        expect *= dense.shape[d]
    if N > 1 and dense.stride(1) != row_elems:
        ok = False
    if B > 1 and dense.stride(0) < N * row_elems:
        ok = False
    if not ok:
        dense = dense.contiguous()
    es = dense.element_size()
    bstride = dense.stride(0) * es if B > 1 else N * row_elems * es
    return dense, row_elems * es, bstride


def _d2j_raw(dense: torch.Tensor, offsets: torch.Tensor, total: Optional[int],
             zero_tail: bool = False) -> torch.Tensor:
    _lib.require_cuda(dense, offsets)
    if dense.dim() < 2:
        raise ValueError("dense_to_jagged: dense tensor must be at least 2-D (B, N, ...)")
    bits = _lib.index_bits(offsets)
    offsets = offsets.contiguous()
    B, N = dense.shape[0], dense.shape[1]
    if offsets.numel() != B + 1:
        raise ValueError(f"dense_to_jagged: offsets has {offsets.numel()} entries, expected B+1={B + 1}")
    if total is None:
        total = int(offsets[-1].item())
    dense, row_bytes, bstride = _dense_layout(dense)
    # zero_tail: `total` may exceed offsets[-1] (rows padded to a fixed bucket); the kernel only
    # writes offsets[-1] rows
    alloc = torch.zeros if zero_tail else torch.empty
    out = alloc((total, *dense.shape[2:]), dtype=dense.dtype, device=dense.device)
    if total > 0:
        _lib.check(_lib.lib().grb_dense_to_jagged(
            dense.data_ptr(), offsets.data_ptr(), out.data_ptr(), B, N, row_bytes, bstride, bits,
            _lib.stream_ptr(dense.device)))
    return out


_PAD_CACHE: dict = {}


def _pad_bytes(value: float, dtype: torch.dtype):
    key = (float(value), dtype)
    buf = _PAD_CACHE.get(key)
    if buf is None:
        raw = torch.tensor([value]).to(dtype).view(torch.uint8).numpy().tobytes()
        buf = C.create_string_buffer(raw, len(raw))
        _PAD_CACHE[key] = buf
    return buf


def _j2d_raw(values: torch.Tensor, offsets: torch.Tensor, N: int, padding_value: float) -> torch.Tensor:
    _lib.require_cuda(values, offsets)
    bits = _lib.index_bits(offsets)
    offsets = offsets.contiguous()
    values = values.contiguous()
    B = offsets.numel() - 1
    row_elems = 1
    for s in values.shape[1:]:
        row_elems *= s
    es = values.element_size()
    out = torch.empty((B, N, *values.shape[1:]), dtype=values.dtype, device=values.device)
    if out.numel() > 0:
        pad = None if padding_value == 0 else _pad_bytes(padding_value, values.dtype)
        _lib.check(_lib.lib().grb_jagged_to_padded_dense(
            values.data_ptr(), offsets.data_ptr(), out.data_ptr(), B, N, row_elems * es, 0,
            C.cast(pad, C.c_void_p) if pad is not None else None, es, bits,
            _lib.stream_ptr(values.device)))
    return out


class _DenseToJagged(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dense, offsets, total, zero_tail=False):
        ctx.save_for_backward(offsets)
        ctx.N = dense.shape[1]
        return _d2j_raw(dense, offsets, total, zero_tail)

    @staticmethod
    def backward(ctx, grad):
        (offsets,) = ctx.saved_tensors
        return _j2d_raw(grad, offsets, ctx.N, 0.0), None, None, None


class _JaggedToPaddedDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, offsets, N, padding_value, padded_rows=False):
        ctx.save_for_backward(offsets)
        ctx.total = values.shape[0]
        ctx.padded_rows = padded_rows
        return _j2d_raw(values, offsets, N, padding_value)

    @staticmethod
    def backward(ctx, grad):
        (offsets,) = ctx.saved_tensors
        return _d2j_raw(grad, offsets, ctx.total, ctx.padded_rows), None, None, None, None


class _GatherLastRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lengths, enc):
        _lib.require_cuda(lengths, enc)
        B, N, D = enc.shape
        enc = enc.contiguous()
        lengths = lengths.contiguous()
        out = torch.empty((B, D), dtype=enc.dtype, device=enc.device)
        if out.numel() > 0:
            _lib.check(_lib.lib().grb_gather_last_rows(
                enc.data_ptr(), lengths.data_ptr(), out.data_ptr(), B, N, D * enc.element_size(),
                _lib.index_bits(lengths), 0, _lib.stream_ptr(enc.device)))
        ctx.save_for_backward(lengths)
        ctx.shape = (B, N, D)
        return out

    @staticmethod
    def backward(ctx, grad):
        (lengths,) = ctx.saved_tensors
        B, N, D = ctx.shape
        grad = grad.contiguous()
        dense = torch.zeros((B, N, D), dtype=grad.dtype, device=grad.device)
        if dense.numel() > 0:
            _lib.check(_lib.lib().grb_gather_last_rows(
                dense.data_ptr(), lengths.data_ptr(), grad.data_ptr(), B, N,
                D * grad.element_size(), _lib.index_bits(lengths), 1,
                _lib.stream_ptr(grad.device)))
        return None, dense


# --------------------------------------------------------------------------------------------
# the nine public functions (reference ops.py signatures)
# --------------------------------------------------------------------------------------------
def asynchronous_complete_cumsum(lengths: torch.Tensor) -> torch.Tensor:
    """(B,) int -> (B+1,) same dtype: [0, cumsum(lengths)].  Reference ops.py:18-38."""
    return _cumsum_raw(lengths)


def dense_to_jagged(dense_tensor: torch.Tensor, offsets: torch.Tensor, *,
                    total: Optional[int] = None, zero_tail: bool = False) -> torch.Tensor:
    """(B, N, ...) -> (offsets[-1], ...): drops the padding.  Reference ops.py:41-64.
    ``total`` (optional) spares the host read of offsets[-1]; with ``zero_tail`` it may be larger
    than offsets[-1] and the extra rows are zero (fixed-size row buckets for CUDA graphs)."""
    if dense_tensor.requires_grad and torch.is_grad_enabled():
        return _DenseToJagged.apply(dense_tensor, offsets, total, zero_tail)
    return _d2j_raw(dense_tensor, offsets, total, zero_tail)


def jagged_to_padded_dense(values: torch.Tensor, offsets: torch.Tensor, max_lengths: int,
                           padding_value: float = 0.0, padded_rows: bool = False) -> torch.Tensor:
    """(T, ...) -> (B, max_lengths, ...) padded with ``padding_value``.  Reference ops.py:67-114
    (same ``ValueError`` for a non-int ``max_lengths``, ops.py:83-84)."""
    if not isinstance(max_lengths, int):
        raise ValueError(f"max_lengths must be an integer, but got {type(max_lengths)}")
    if values.requires_grad and torch.is_grad_enabled():
        return _JaggedToPaddedDense.apply(values, offsets, max_lengths, padding_value, padded_rows)
    return _j2d_raw(values, offsets, max_lengths, padding_value)


def batch_gather_embeddings(rowwise_indices: torch.Tensor, embeddings: torch.Tensor) -> torch.Tensor:
    """(B, N) indices into (B, X, D) -> (B, N, D).  Reference ops.py:117-146."""
    D = embeddings.size(-1)
    idx = rowwise_indices.unsqueeze(-1).expand(-1, -1, D)
    return torch.gather(embeddings, 1, idx)


def batch_scatter_embeddings(dst_embeddings: torch.Tensor, rowwise_indices: torch.Tensor,
                             src_embeddings: torch.Tensor) -> None:
    """dst[b, idx[b]] = src[b] in place.  Reference ops.py:149-168."""
    D = dst_embeddings.size(-1)
    idx = rowwise_indices.view(-1, 1, 1).expand(-1, 1, D)
    dst_embeddings.scatter_(1, idx, src_embeddings.unsqueeze(1))


def get_current_embeddings(lengths: torch.Tensor, encoded_embeddings: torch.Tensor) -> torch.Tensor:
    """(B, N, D) -> (B, D): row lengths[b]-1 of each sequence.  Reference ops.py:171-187."""
    return _GatherLastRows.apply(lengths, encoded_embeddings)


def jagged_or_dense_repeat_interleave_dim0(x: torch.Tensor, lengths: torch.Tensor,
                                           repeats: int) -> torch.Tensor:
    """Reference ops.py:190-207."""
    if x.dim() == 3:
        return x.repeat_interleave(repeats, dim=0)
    assert x.dim() == 2, f"x.size() = {x.size()}"
    padded = jagged_to_padded_dense(x, asynchronous_complete_cumsum(lengths),
                                    int(lengths.max().item()), 0.0)
    rep_lengths = lengths.repeat_interleave(repeats, dim=0)
    return dense_to_jagged(padded.repeat_interleave(repeats, dim=0),
                           asynchronous_complete_cumsum(rep_lengths))


def jagged_or_dense_index_select_dim0(x: torch.Tensor, lengths: torch.Tensor,
                                      indices: torch.Tensor) -> torch.Tensor:
    """Reference ops.py:210-226."""
    if x.dim() == 3:
        return x[indices, :, :]
    assert x.dim() == 2, f"x.size() = {x.size()}"
    padded = jagged_to_padded_dense(x, asynchronous_complete_cumsum(lengths),
                                    int(lengths.max().item()), 0.0)
    return dense_to_jagged(padded[indices, :], asynchronous_complete_cumsum(lengths[indices]))


def mask_dense_by_aux_mask(dense_tensor: torch.Tensor, aux_mask: torch.Tensor,
                           lengths: torch.Tensor, max_lengths: int):
    """Keep the rows selected by ``aux_mask`` (B, N) and re-pad.  Reference ops.py:229-260."""
    offsets = asynchronous_complete_cumsum(lengths)
    total = int(offsets[-1].item())
    jagged = dense_to_jagged(dense_tensor, offsets, total=total)
    jagged_mask = dense_to_jagged(aux_mask, offsets, total=total)
    kept = jagged[jagged_mask]
    new_lengths = aux_mask.int().sum(dim=1)
    return jagged_to_padded_dense(kept, asynchronous_complete_cumsum(new_lengths), max_lengths,
                                  0.0), new_lengths
