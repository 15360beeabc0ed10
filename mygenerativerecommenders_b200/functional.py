"""Autograd-aware wrappers over the C ABI for the fused hot-path kernels.

Each function documents the reference lines it replaces (paths under
/root/reference/src/generative_recommenders_pl/models/).  CUDA tensors only; errors from the
library surface as Python exceptions (ValueError / NotImplementedError / RuntimeError).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch

from . import _lib


def _ld(t: torch.Tensor) -> int:
    """Row stride (elements) of a 2-D tensor whose rows are contiguous."""
    return t.stride(0) if t.shape[0] > 1 else t.shape[1]


def _rows_contiguous(t: torch.Tensor) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        return t.contiguous()
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        return t.contiguous()
    return t


# --------------------------------------------------------------------------------------------
# fused jagged HSTU attention  (sequential_encoders/hstu.py:96-128 + :134-205)
# --------------------------------------------------------------------------------------------
def _attn_args(q, k, v, offsets, timestamps, ts_w, pos_w, thresholds, N, H, dqk, dv, max_len,
               cache=None, short=False):
    a = _lib.HstuAttnArgs()
    a.B = offsets.numel() - 1
    a.N = N
    a.T = q.shape[0]
    a.max_len = max_len
    a.H, a.dqk, a.dv = H, dqk, dv
    a.dtype = _lib.dtype_code(q.dtype)
    a.index_bits = _lib.index_bits(offsets)
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ldq, a.ldk, a.ldv = _ld(q), _ld(k), _ld(v)
    a.offsets = offsets.data_ptr()
    if timestamps is not None:
        a.num_buckets = thresholds.numel()
        a.timestamps = timestamps.data_ptr()
        a.ts_w = ts_w.data_ptr()
        a.pos_w = pos_w.data_ptr()
        a.bucket_thresholds = thresholds.data_ptr()
        oct_t = bucket_octaves(thresholds)
        a.bucket_octaves = oct_t.data_ptr()
        if cache is not None and not short:
            a.bucket_cache = cache.data_ptr()
            a.bucket_cache_max_len = cache.grb_max_len
    if short:     # masked bucket tiles + item schedule: the short-sequence kernels (with or without bias)
        a.bucket_cache = cache.data_ptr()
        a.bucket_cache_max_len = cache.grb_max_len
        a.bucket_cache_masked = 1
        a.short_schedule = cache.grb_sched.data_ptr()
    return a


def bucket_octaves(thresholds: torch.Tensor) -> torch.Tensor:
    """Device copy of grb_bucket_octaves(thresholds).  The table rides on the thresholds tensor object
    itself (the module's ``_bucket_thresholds`` buffer), so it lives exactly as long as the thresholds
    it was computed from: no global cache keyed by a device address that the allocator can hand to
    another module's table.  The one synchronous ``.cpu()`` happens on the first use of a buffer, i.e.
    in the warm-up iterations before any CUDA-graph capture."""
    hit = getattr(thresholds, "_grb_octaves", None)
    if hit is not None and hit[0] == thresholds._version and hit[1].device == thresholds.device:
        return hit[1]
    host = thresholds.detach().cpu().contiguous()
    out = torch.empty(130, dtype=torch.int32)
    _lib.check(_lib.lib().grb_bucket_octaves(host.data_ptr(), host.numel(), out.data_ptr()))
    dev = out.to(thresholds.device)
    thresholds._grb_octaves = (thresholds._version, dev)
    return dev


class BucketCache(torch.Tensor):
    """uint8 tensor holding grb_hstu_bucket_tiles output; remembers the max_len it was built for.
    ``grb_masked``: built by grb_hstu_bucket_tiles_masked for the short-sequence kernels (masked pairs
    hold 255), with the item schedule of the batch in ``grb_sched``."""
    grb_max_len: int = 0
    grb_masked: bool = False
    grb_sched: Optional[torch.Tensor] = None


def hstu_bucket_cache(offsets: torch.Tensor, timestamps: Optional[torch.Tensor],
                      thresholds: Optional[torch.Tensor], N: int, max_len: Optional[int] = None,
                      masked: bool = False) -> torch.Tensor:
    """Tabulate bucket(|ts[b, i+1] - ts[b, j]|) for every causal 128x128 tile of every sequence,
    once per batch (the reference recomputes it per layer, hstu.py:117-123).  Pass the result as
    ``bucket_cache=`` to :func:`hstu_attention` for all layers, forward and backward.

    ``masked`` (short sequences, max_len <= 256): the tiles also carry the causal / length masks
    (bucket 255) and the batch's item schedule is attached; ``timestamps`` may then be None (no
    relative bias: masks only)."""
    _lib.require_cuda(offsets, timestamps, thresholds)
    max_len = N if max_len is None else min(max_len, N)
    B = offsets.numel() - 1
    nbytes = int(_lib.lib().grb_hstu_bucket_cache_bytes(B, max_len))
    cache = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=offsets.device).as_subclass(BucketCache)
    cache.grb_max_len = max_len
    cache.grb_masked = bool(masked)
    offsets = offsets.contiguous()
    stream = _lib.stream_ptr(offsets.device)
    if timestamps is not None:
        timestamps = timestamps.contiguous()
        ts_args = (timestamps.data_ptr(), B, N, max_len, thresholds.data_ptr(), thresholds.numel(),
                   bucket_octaves(thresholds).data_ptr())
    else:
        if not masked:
            raise ValueError("hstu_bucket_cache: timestamps are required unless masked=True")
        ts_args = (None, B, N, max_len, None, 0, None)
    fn = _lib.lib().grb_hstu_bucket_tiles_masked if masked else _lib.lib().grb_hstu_bucket_tiles
    with _lib.timed("hstu_bucket_tiles"):
        _lib.check(fn(offsets.data_ptr(), _lib.index_bits(offsets), *ts_args, cache.data_ptr(), stream))
    if masked:
        sched = torch.empty(B + 2, dtype=torch.int32, device=offsets.device)
        _lib.check(_lib.lib().grb_hstu_short_schedule(offsets.data_ptr(), _lib.index_bits(offsets), B, N,
                                                      sched.data_ptr(), stream))
        cache.grb_sched = sched
    return cache


SHORT_MAX_LEN = 256


def short_path_applies(q: torch.Tensor, dqk: int, dv: int, max_len: int) -> bool:
    """The short-sequence tcgen05 kernels (csrc/hstu_attn_short.cu): bf16, 64-wide heads,
    every sequence <= 256 tokens.  GRB_NO_SHORT=1 keeps the long-sequence kernels (A/B switch)."""
    import os
    return (q.dtype == torch.bfloat16 and dqk == 64 and dv == 64 and 0 < max_len <= SHORT_MAX_LEN
            and os.environ.get("GRB_NO_SHORT") != "1")


def _zero_tail_rows(t: torch.Tensor, n_mats: int, offsets: torch.Tensor) -> None:
    """Rows >= offsets[-1] of the n_mats stacked (rows, W) matrices in ``t`` := 0 (fixed row buckets: the
    padding rows are the only ones the attention kernels leave unwritten)."""
    rows, W = t.shape[-2], t.shape[-1]
    row_bytes = W * t.element_size()
    if row_bytes % 16 or t.data_ptr() % 16:
        t.zero_()
        return
    _lib.check(_lib.lib().grb_zero_tail_rows(
        t.data_ptr(), row_bytes, rows, row_bytes, n_mats, rows * row_bytes, offsets.data_ptr(),
        _lib.index_bits(offsets), offsets.numel() - 1, rows, _lib.stream_ptr(t.device)))


class _HstuAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, offsets, timestamps, ts_w, pos_w, thresholds, N, H, dqk, dv, max_len,
                cache=None, rows_padded=False):
        _lib.require_cuda(q, k, v, offsets, timestamps, ts_w, pos_w, thresholds)
        if not (q.dtype == k.dtype == v.dtype):
            raise ValueError("hstu_attention: q, k, v must share a dtype")
        q, k, v = _rows_contiguous(q), _rows_contiguous(k), _rows_contiguous(v)
        offsets = offsets.contiguous()
        if timestamps is not None:
            if timestamps.dtype != torch.int64 or timestamps.shape != (offsets.numel() - 1, N):
                raise ValueError("hstu_attention: timestamps must be int64 of shape (B, N)")
            timestamps = timestamps.contiguous()
            ts_w = ts_w.detach().float().contiguous()
            pos_w = pos_w.detach().float().contiguous()
            if ts_w.numel() != thresholds.numel() + 1 or pos_w.numel() < 2 * N - 1:
                raise ValueError("hstu_attention: bias table sizes do not match N / num_buckets")
        # rows_padded: q/k/v carry rows past offsets[-1] (fixed-size row buckets); the kernels
        # never write those rows, so they must start as zeros
        out = torch.empty((q.shape[0], H * dv), dtype=q.dtype, device=q.device)
        if cache is not None and ((timestamps is None and not cache.grb_masked) or cache.grb_max_len != max_len):
            cache = None
        short = short_path_applies(q, dqk, dv, max_len) and (timestamps is None or thresholds.numel() <= 254)
        if short:
            if cache is None or not cache.grb_masked:
                cache = hstu_bucket_cache(offsets, timestamps, thresholds, N, max_len, masked=True)
        elif cache is not None and cache.grb_masked:
            cache = None            # masked tiles only serve the short-sequence kernels
        a = _attn_args(q, k, v, offsets, timestamps, ts_w, pos_w, thresholds, N, H, dqk, dv, max_len, cache,
                       short)
        a.out, a.ldo = out.data_ptr(), H * dv
        if rows_padded:
            # the short-sequence kernels zero the rows past offsets[-1] themselves
            if short and out.data_ptr() % 16 == 0 and (H * dv * out.element_size()) % 16 == 0:
                a.zero_tail_rows = 1
            else:
                _zero_tail_rows(out, 1, offsets)
        with _lib.timed("hstu_attn_fwd"):
            _lib.check(_lib.lib().grb_hstu_attn_fwd(C.byref(a), _lib.stream_ptr(q.device)))
        ctx.save_for_backward(q, k, v, offsets, timestamps, ts_w, pos_w, thresholds)
        ctx.dims = (N, H, dqk, dv, max_len)
        ctx.cache = cache
        ctx.short = short
        ctx.rows_padded = rows_padded
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, offsets, timestamps, ts_w, pos_w, thresholds = ctx.saved_tensors
        N, H, dqk, dv, max_len = ctx.dims
        dout = _rows_contiguous(dout)
        T = q.shape[0]
        # one allocation (and, with padded rows, one fill) for the three gradients; one zero-filled
        # fp32 workspace for the dQ accumulator and the privatised bias-gradient copies
        if dqk == dv:
            g3 = torch.empty((3, T, H * dqk), dtype=q.dtype, device=q.device)
            dq, dk, dvv = g3[0], g3[1], g3[2]
            zero_in_kernel = (ctx.rows_padded and ctx.short and g3.data_ptr() % 16 == 0
                              and (H * dqk * g3.element_size()) % 16 == 0 and (T * H * dqk * g3.element_size()) % 16 == 0)
            if ctx.rows_padded and not zero_in_kernel:
                _zero_tail_rows(g3, 3, offsets)
        else:
            dq = torch.empty((T, H * dqk), dtype=q.dtype, device=q.device)
            dk = torch.empty((T, H * dqk), dtype=q.dtype, device=q.device)
            dvv = torch.empty((T, H * dv), dtype=q.dtype, device=q.device)
            zero_in_kernel = False
            if ctx.rows_padded:
                for t in (dq, dk, dvv):
                    _zero_tail_rows(t, 1, offsets)
        short = ctx.short
        # long-sequence kernels: fp32 dQ accumulator (zero-filled).  Short kernels: plain scratch for
        # the partial dQ of the second query tile, only when a sequence can have one.
        n_acc = T * H * dqk if (not short or max_len > 128) else 0
        copies = n_ts = n_pos = 0
        if timestamps is not None:
            # every CTA adds into d ts_w / d pos_w; with short sequences that is thousands of CTAs
            # on a handful of cache lines.  Give them private copies (<= 256) and sum after.
            # (The short-sequence kernels are persistent: one copy per CTA is 2 x SMs at most.)
            copies = 64 if short else max(1, min(256, (4 << 20) // max(1, pos_w.numel())))
            n_ts, n_pos = ts_w.numel(), pos_w.numel()
        if short:
            ws = torch.empty(n_acc + copies * (n_ts + n_pos), dtype=torch.float32, device=q.device)
            ws[n_acc:].zero_()
        else:
            ws = torch.zeros(n_acc + copies * (n_ts + n_pos), dtype=torch.float32, device=q.device)
        dq_acc = ws[:n_acc]
        a = _attn_args(q, k, v, offsets, timestamps, ts_w, pos_w, thresholds, N, H, dqk, dv, max_len,
                       ctx.cache, short)
        a.dout, a.lddo = dout.data_ptr(), _ld(dout)
        a.dq, a.dk, a.dv_grad = dq.data_ptr(), dk.data_ptr(), dvv.data_ptr()
        a.zero_tail_rows = 1 if zero_in_kernel else 0
        a.lddq, a.lddk, a.lddv = H * dqk, H * dqk, H * dv
        a.dq_accum = dq_acc.data_ptr() if n_acc else None
        d_ts = d_pos = None
        if timestamps is not None:
            d_ts = ws[n_acc:n_acc + copies * n_ts].view(copies, n_ts)
            d_pos = ws[n_acc + copies * n_ts:].view(copies, n_pos)
            a.d_ts_w, a.d_pos_w, a.d_bias_copies = d_ts.data_ptr(), d_pos.data_ptr(), copies
        with _lib.timed("hstu_attn_bwd"):
            _lib.check(_lib.lib().grb_hstu_attn_bwd(C.byref(a), _lib.stream_ptr(q.device)))
        if d_pos is not None:
            if copies > 1:     # both tables' copies summed by one launch (ATen: two single-block reductions)
                sums = torch.empty(n_ts + n_pos, dtype=torch.float32, device=q.device)
                _lib.check(_lib.lib().grb_colsum_f32_pair(
                    d_ts.data_ptr(), n_ts, sums.data_ptr(), d_pos.data_ptr(), n_pos,
                    sums.data_ptr() + 4 * n_ts, copies, _lib.stream_ptr(q.device)))
                d_ts, d_pos = sums[:n_ts], sums[n_ts:]
            else:
                d_pos, d_ts = d_pos[0], d_ts[0]
        return dq, dk, dvv, None, None, d_ts, d_pos, None, None, None, None, None, None, None, None


def hstu_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, offsets: torch.Tensor,
                   timestamps: Optional[torch.Tensor], ts_w: Optional[torch.Tensor],
                   pos_w: Optional[torch.Tensor], bucket_thresholds: Optional[torch.Tensor],
                   N: int, num_heads: int, attention_dim: int, linear_dim: int,
                   max_len: Optional[int] = None,
                   bucket_cache: Optional[torch.Tensor] = None,
                   rows_padded: bool = False) -> torch.Tensor:
    """Jagged pointwise-SiLU attention with relative time+position bias.

    q, k: (T, H*attention_dim); v: (T, H*linear_dim); offsets (B+1); timestamps (B, N) int64 or
    None (no bias, hstu.py:191).  Returns (T, H*linear_dim).  ``N`` is the padded length that the
    reference uses for the 1/N scale and the bias origin (hstu.py:150,193)."""
    if max_len is None:
        max_len = N
    if ts_w is not None and ts_w.dtype != torch.float32:
        ts_w = ts_w.float()
    if pos_w is not None and pos_w.dtype != torch.float32:
        pos_w = pos_w.float()
    return _HstuAttention.apply(q, k, v, offsets, timestamps, ts_w, pos_w, bucket_thresholds,
                                N, num_heads, attention_dim, linear_dim, min(max_len, N),
                                bucket_cache, rows_padded)


def hstu_attention_decode(q: torch.Tensor, k_cache: torch.Tensor, v: torch.Tensor,
                          offsets: torch.Tensor, positions: torch.Tensor,
                          timestamps: Optional[torch.Tensor], ts_w: Optional[torch.Tensor],
                          pos_w: Optional[torch.Tensor], bucket_thresholds: Optional[torch.Tensor],
                          N: int, num_heads: int, attention_dim: int, linear_dim: int) -> torch.Tensor:
    """The attention row of ONE position per sequence, from the caches of the incremental path
    (hstu.py:151-177 + :179-204 + :397-401: the reference recomputes the whole padded attention and
    keeps rows ``delta_x_offsets[0]``; row p only depends on keys / values 0..p).

    q (B, H*attention_dim): the new query rows; k_cache (B, N, H*attention_dim): padded keys, already
    holding the new key at [b, positions[b]]; v (T, H*linear_dim): jagged values, new row included;
    positions (B) int32 | int64 = delta_x_offsets[1].  Returns (B, H*linear_dim).  Inference only."""
    _lib.require_cuda(q, k_cache, v, offsets, positions, timestamps)
    if torch.is_grad_enabled() and (q.requires_grad or k_cache.requires_grad or v.requires_grad):
        raise NotImplementedError(
            "hstu_attention_decode is forward-only (the incremental path serves inference); "
            "run it under torch.no_grad() / inference_mode()")
    B = offsets.numel() - 1
    if q.dtype != k_cache.dtype or q.dtype != v.dtype:
        raise ValueError("hstu_attention_decode: q, k_cache and v must share a dtype")
    q = _rows_contiguous(q)
    v = _rows_contiguous(v)
    kc = k_cache.reshape(B * N, -1) if k_cache.is_contiguous() else k_cache.contiguous().view(B * N, -1)
    positions = positions.contiguous()
    out = torch.empty((B, num_heads * linear_dim), dtype=q.dtype, device=q.device)
    a = _lib.HstuAttnDecodeArgs()
    a.B, a.N, a.H, a.dqk, a.dv = B, N, num_heads, attention_dim, linear_dim
    a.dtype = _lib.dtype_code(q.dtype)
    a.index_bits, a.pos_bits = _lib.index_bits(offsets), _lib.index_bits(positions)
    a.q, a.ldq = q.data_ptr(), _ld(q)
    a.k_cache, a.ldk = kc.data_ptr(), kc.stride(0)
    a.v, a.ldv = v.data_ptr(), _ld(v)
    a.offsets, a.positions = offsets.data_ptr(), positions.data_ptr()
    if timestamps is not None:
        ts_w = ts_w if ts_w.dtype == torch.float32 else ts_w.float()
        pos_w = pos_w if pos_w.dtype == torch.float32 else pos_w.float()
        a.num_buckets = bucket_thresholds.numel()
        a.timestamps, a.ts_w, a.pos_w = timestamps.data_ptr(), ts_w.data_ptr(), pos_w.data_ptr()
        a.bucket_thresholds = bucket_thresholds.data_ptr()
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    with _lib.timed("hstu_attn_decode"):
        _lib.check(_lib.lib().grb_hstu_attn_decode(C.byref(a), _lib.stream_ptr(q.device)))
    return out


# --------------------------------------------------------------------------------------------
# y = gate * LayerNorm(x)   (hstu.py:258-264, :300, :402)
# --------------------------------------------------------------------------------------------
def _ln_args(x, gate, mean, rstd, eps, p_drop, seed, salt):
    a = _lib.LnGateArgs()
    a.x, a.ldx = x.data_ptr(), _ld(x)
    if gate is not None:
        a.gate, a.ldg = gate.data_ptr(), _ld(gate)
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    a.rows, a.W = x.shape
    a.eps, a.dtype = float(eps), _lib.dtype_code(x.dtype)
    if p_drop > 0.0:
        a.p_drop, a.seed, a.salt = float(p_drop), seed.data_ptr(), int(salt)
    return a


class _LnGate(torch.autograd.Function):
    """y = dropout(gate * LN(x)); with ``skip`` a second output aliases x (the residual operand of the
    layer, hstu.py:413), so that both gradients of x meet in ONE backward call and leave it as one tensor:
    dx = LN backward + d skip, summed inside the kernel instead of by autograd's own add."""

    @staticmethod
    def forward(ctx, x, gate, eps, p_drop, seed, salt, skip):
        _lib.require_cuda(x, gate, seed)
        ctx.set_materialize_grads(False)     # an unused output's gradient arrives as None, not as zeros
        x = _rows_contiguous(x)
        if gate is not None:
            gate = _rows_contiguous(gate)
            if gate.dtype != x.dtype or gate.shape != x.shape:
                raise ValueError("ln_gate: gate must match x in shape and dtype")
        rows, W = x.shape
        y = torch.empty((rows, W), dtype=x.dtype, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        a = _ln_args(x, gate, mean, rstd, eps, p_drop, seed, salt)
        a.y, a.ldy = y.data_ptr(), W
        _lib.check(_lib.lib().grb_ln_gate_fwd_ex(C.byref(a), _lib.stream_ptr(x.device)))
        ctx.save_for_backward(x, gate, mean, rstd, seed)
        ctx.cfg = (float(eps), float(p_drop), int(salt))
        if skip:
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x, gate, mean, rstd, seed = ctx.saved_tensors
        eps, p_drop, salt = ctx.cfg
        rows, W = x.shape
        if dy is None:                       # only the residual branch carried a gradient
            return dskip, None, None, None, None, None, None
        dy = _rows_contiguous(dy)
        dx = torch.empty((rows, W), dtype=x.dtype, device=x.device)
        dgate = torch.empty((rows, W), dtype=x.dtype, device=x.device) if gate is not None else None
        a = _ln_args(x, gate, mean, rstd, eps, p_drop, seed, salt)
        a.y, a.ldy = dy.data_ptr(), _ld(dy)
        a.dx, a.lddx = dx.data_ptr(), W
        if dgate is not None:
            a.dgate, a.lddg = dgate.data_ptr(), W
        if dskip is not None:
            dskip = _rows_contiguous(dskip if dskip.dtype == x.dtype else dskip.to(x.dtype))
            a.res, a.ldres = dskip.data_ptr(), _ld(dskip)
        _lib.check(_lib.lib().grb_ln_gate_bwd_ex(C.byref(a), _lib.stream_ptr(x.device)))
        return dx, dgate, None, None, None, None, None


def ln_dropout_fusable(x: torch.Tensor) -> bool:
    """Whether ``layer_norm_gate`` can draw the dropout mask inside its kernels for rows like x's."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.shape[1] in (256, 512, 1024)
            and x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0)


class _SiluSplit(torch.autograd.Function):
    """torch.split(F.silu(x), sizes, dim=1) (hstu.py:304-320).  Backward: the pieces' gradients come
    back separately; one kernel reads them in place and applies silu' (autograd: cat + silu_backward,
    two full passes)."""

    @staticmethod
    def forward(ctx, x, *sizes):
        _lib.require_cuda(x)
        x = _rows_contiguous(x)
        rows, W = x.shape
        y = torch.empty((rows, W), dtype=x.dtype, device=x.device)
        _lib.check(_lib.lib().grb_silu_fwd(x.data_ptr(), _ld(x), y.data_ptr(), W, rows, W,
                                           _lib.dtype_code(x.dtype), _lib.stream_ptr(x.device)))
        ctx.save_for_backward(x)
        ctx.sizes = sizes
        return tuple(y.split(list(sizes), dim=1))

    @staticmethod
    def backward(ctx, *grads):
        (x,) = ctx.saved_tensors
        rows, W = x.shape
        n = len(ctx.sizes)
        es = x.element_size()

        def prep(g):
            if g is None:
                return None
            g = _rows_contiguous(g if g.dtype == x.dtype else g.to(x.dtype))
            if g.data_ptr() % 16 or (_ld(g) * es) % 16:      # the kernel reads 16-byte chunks
                g = g.clone(memory_format=torch.contiguous_format)
            return g

        keep = [prep(g) for g in grads]
        dx = torch.empty((rows, W), dtype=x.dtype, device=x.device)
        ptrs = (C.c_void_p * n)(*[None if g is None else g.data_ptr() for g in keep])
        lds = (C.c_int64 * n)(*[0 if g is None else _ld(g) for g in keep])
        widths = (C.c_int32 * n)(*ctx.sizes)
        _lib.check(_lib.lib().grb_silu_split_bwd(
            x.data_ptr(), _ld(x), n, ptrs, lds, widths, dx.data_ptr(), W, rows,
            _lib.dtype_code(x.dtype), _lib.stream_ptr(x.device)))
        return (dx,) + (None,) * n


def silu_split(x: torch.Tensor, sizes) -> Tuple[torch.Tensor, ...]:
    """``torch.split(F.silu(x), sizes, dim=1)`` for a 2-D activation with up to four column blocks
    whose widths are multiples of 16 bytes (the u, v, q, k blocks of the HSTU layer)."""
    _lib.require_cuda(x)
    sizes = [int(v) for v in sizes]
    per = 16 // x.element_size()
    if (x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16) and 1 <= len(sizes) <= 4
            and sum(sizes) == x.shape[1] and all(v > 0 and v % per == 0 for v in sizes)):
        return _SiluSplit.apply(x, *sizes)
    return torch.split(torch.nn.functional.silu(x), sizes, dim=1)


class _LinearBias(torch.autograd.Function):
    """F.linear(x, w, b) for 2-D x whose bias gradient is a (1, T) x (T, out) GEMM instead of
    ATen's column reduction (25 us for a 14k x 256 bf16 gradient, 4 per step at the C2 shape)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return torch.addmm(b, x, w.t())

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.mm(g, w) if ctx.needs_input_grad[0] else None
        dw = torch.mm(g.t(), x) if ctx.needs_input_grad[1] else None
        db = torch.mm(g.new_ones(1, g.shape[0]), g).view(-1) if ctx.needs_input_grad[2] else None
        return dx, dw, db


def _mm_out(a: torch.Tensor, b: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """a @ b with the result written in out_dtype by the GEMM itself (cuBLAS fp32 epilogue)."""
    if out_dtype == a.dtype:
        return torch.mm(a, b)
    return torch.mm(a, b, out_dtype=out_dtype)


class _MasterLinear(torch.autograd.Function):
    """Projection of low-precision activations by fp32 master parameters (the two projections of
    the HSTU layer, hstu.py:302-304 and :404-413, under bf16 compute).  The parameters are cast to
    the activation dtype inside; backward lets the GEMMs write the weight and bias gradients in
    the parameters' dtype directly, instead of a bf16 gradient that autograd then casts (one
    elementwise launch per parameter and step, and a rounding that the fp32 accumulator did not need).

    ``w_in_out``: w is (in, out) and y = x @ w (``_uvqk``); else w is (out, in) and
    y = x @ w.T + b (``nn.Linear``)."""

    @staticmethod
    def forward(ctx, x, w, b, w_in_out):
        wc = w if w.dtype == x.dtype else w.to(x.dtype)
        ctx.save_for_backward(x, wc)
        ctx.w_in_out = w_in_out
        ctx.w_dtype = w.dtype
        ctx.b_dtype = b.dtype if b is not None else None
        if w_in_out or b is None:
            return torch.mm(x, wc if w_in_out else wc.t())
        bc = b if b.dtype == x.dtype else b.to(x.dtype)
        return torch.addmm(bc, x, wc.t())

    @staticmethod
    def backward(ctx, g):
        x, wc = ctx.saved_tensors
        g = g.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.mm(g, wc.t() if ctx.w_in_out else wc)
        if ctx.needs_input_grad[1]:
            dw = _mm_out(x.t(), g, ctx.w_dtype) if ctx.w_in_out else _mm_out(g.t(), x, ctx.w_dtype)
        if ctx.b_dtype is not None and ctx.needs_input_grad[2]:
            db = _mm_out(g.new_ones(1, g.shape[0]), g, ctx.b_dtype).view(-1)
        return dx, dw, db, None


def master_linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None,
                  w_in_out: bool = False) -> torch.Tensor:
    """x @ w (``w_in_out``) or F.linear(x, w, b), with w / b possibly fp32 masters of a bf16 x."""
    _lib.require_cuda(x, w, b)
    if x.dim() == 2:
        return _MasterLinear.apply(x, w, b, w_in_out)
    wc = w.to(x.dtype)
    if w_in_out:
        return torch.mm(x, wc)
    return torch.nn.functional.linear(x, wc, None if b is None else b.to(x.dtype))


def linear_bias(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """x @ w.T + b (hstu.py:404-413 output projection)."""
    _lib.require_cuda(x, w, b)
    if x.dim() == 2 and b is not None:
        return _LinearBias.apply(x, w, b)
    return torch.nn.functional.linear(x, w, b)


# --------------------------------------------------------------------------------------------
# f1: the layer's projections on the tcgen05 GEMM with fused epilogues (csrc/proj_gemm.cu)
#     hstu.py:302-320 (UVQK + SiLU + split) and :404-413 (_o + bias + residual)
# --------------------------------------------------------------------------------------------
def _proj_gemm(A, B, M, N, K, a_mn, b_mn, epi, out0, out1=None, bias=None, res=None):
    a = _lib.ProjGemmArgs()
    a.M, a.N, a.K = M, N, K
    a.a_mn, a.b_mn, a.epi = int(a_mn), int(b_mn), epi
    a.A, a.lda = A.data_ptr(), A.stride(0)
    a.B, a.ldb = B.data_ptr(), B.stride(0)
    a.out0, a.ldo0 = out0.data_ptr(), out0.stride(0)
    if out1 is not None:
        a.out1, a.ldo1 = out1.data_ptr(), out1.stride(0)
    if bias is not None:
        a.bias = bias.data_ptr()
    if res is not None:
        a.res, a.ldres = res.data_ptr(), res.stride(0)
    _lib.check(_lib.lib().grb_proj_gemm(C.byref(a), _lib.stream_ptr(A.device)))


def _gemm_ok(x: torch.Tensor, K: int, N: int) -> bool:
    """bf16 rows on the GPU, shapes the 128 x 256 x 64 tiles take.  GRB_NO_PROJ_GEMM=1 keeps the
    cuBLAS composite (A/B switch)."""
    import os
    return (x.is_cuda and x.dim() == 2 and x.dtype == torch.bfloat16 and K % 64 == 0 and N % 256 == 0
            and x.shape[0] > 0 and os.environ.get("GRB_NO_PROJ_GEMM") != "1")


def _al16(t: torch.Tensor) -> torch.Tensor:
    """Rows contiguous, 16-byte aligned base and row stride (what TMA needs)."""
    t = _rows_contiguous(t)
    if t.data_ptr() % 16 or (t.stride(0) * t.element_size()) % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _silu_split_bwd(x_pre: torch.Tensor, grads, sizes) -> torch.Tensor:
    """d(pre-activation) from the gradients of the column blocks of SiLU(x_pre): one pass."""
    rows, W = x_pre.shape
    n = len(sizes)
    es = x_pre.element_size()

    def prep(g):
        if g is None:
            return None
        g = _rows_contiguous(g if g.dtype == x_pre.dtype else g.to(x_pre.dtype))
        if g.data_ptr() % 16 or (_ld(g) * es) % 16:      # the kernel reads 16-byte chunks
            g = g.clone(memory_format=torch.contiguous_format)
        return g

    keep = [prep(g) for g in grads]
    dx = torch.empty((rows, W), dtype=x_pre.dtype, device=x_pre.device)
    ptrs = (C.c_void_p * n)(*[None if g is None else g.data_ptr() for g in keep])
    lds = (C.c_int64 * n)(*[0 if g is None else _ld(g) for g in keep])
    widths = (C.c_int32 * n)(*sizes)
    _lib.check(_lib.lib().grb_silu_split_bwd(
        x_pre.data_ptr(), _ld(x_pre), n, ptrs, lds, widths, dx.data_ptr(), W, rows,
        _lib.dtype_code(x_pre.dtype), _lib.stream_ptr(x_pre.device)))
    return dx


def cast_many_bf16(tensors) -> list:
    """bf16 copies of a list of fp32 CUDA tensors, one launch (``grb_cast_f32_bf16_many``), carved out of
    one buffer (every copy 16-byte aligned).  Not differentiable: the copies are operands, the gradients
    go to the fp32 masters."""
    tensors = [t.detach() for t in tensors]
    if not tensors:
        return []
    _lib.require_cuda(*tensors)
    for t in tensors:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("cast_many_bf16: contiguous float32 tensors only")
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += (t.numel() + 7) // 8 * 8
    flat = torch.empty(total, dtype=torch.bfloat16, device=tensors[0].device)
    outs = [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, tensors)]
    n = len(tensors)
    src = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    dst = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    num = (C.c_int64 * n)(*[t.numel() for t in tensors])
    _lib.check(_lib.lib().grb_cast_f32_bf16_many(n, src, dst, num, _lib.stream_ptr(flat.device)))
    return outs


def zeros_many(shapes, device) -> list:
    """fp32 zero tensors of the given shapes carved out of ONE zero-filled buffer (one fill launch; every
    tensor 256-byte aligned): the split-K weight-gradient accumulators of all layers of a step."""
    offs, total = [], 0
    for sh in shapes:
        offs.append(total)
        total += (math.prod(sh) + 63) // 64 * 64
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    return [flat[o:o + math.prod(sh)].view(sh) for o, sh in zip(offs, shapes)]


def _usable_shadow(w: torch.Tensor, w_cast, dtype) -> bool:
    return (w_cast is not None and w_cast.dtype == dtype and w_cast.shape == w.shape and w_cast.is_contiguous()
            and w_cast.device == w.device and w_cast.data_ptr() % 16 == 0)


def _take_buf(ctx, name: str, shape, device) -> torch.Tensor:
    """A pre-zeroed accumulator handed in by the caller (used once: a second backward through the same
    node gets a fresh one), else a new zero tensor."""
    buf = getattr(ctx, name, None)
    setattr(ctx, name, None)
    if buf is not None and tuple(buf.shape) == tuple(shape) and buf.dtype == torch.float32 and buf.device == device:
        return buf
    return torch.zeros(shape, dtype=torch.float32, device=device)


class _UvqkProj(torch.autograd.Function):
    """split(SiLU(xn @ W_uvqk)) (hstu.py:302-320).  Forward: one GEMM whose epilogue writes the
    pre-activation and its SiLU.  Backward: SiLU' over the four consumers' gradients (one pass), the
    input gradient and the fp32 weight gradient (split-K over the token dimension) on the same kernel."""

    @staticmethod
    def forward(ctx, xn, w, w_cast, dw_buf, *sizes):
        xn = _al16(xn)
        T, D = xn.shape
        Ntot = w.shape[1]
        if _usable_shadow(w, w_cast, xn.dtype):
            wc = w_cast
        else:
            wc = (w if w.dtype == xn.dtype else w.to(xn.dtype)).contiguous()   # (K = D, N = Ntot) row-major
        ctx.dw_buf = dw_buf
        pre = torch.empty((T, Ntot), dtype=xn.dtype, device=xn.device)
        act = torch.empty((T, Ntot), dtype=xn.dtype, device=xn.device)
        with _lib.timed("proj_gemm_uvqk_fwd"):
            _proj_gemm(xn, wc, T, Ntot, D, False, True, _lib.GEMM_EPI_SILU2, pre, act)
        ctx.save_for_backward(xn, wc, pre)
        ctx.sizes = sizes
        ctx.w_dtype = w.dtype
        return tuple(act.split(list(sizes), dim=1))

    @staticmethod
    def backward(ctx, *grads):
        xn, wc, pre = ctx.saved_tensors
        T, D = xn.shape
        Ntot = wc.shape[1]
        dx = _silu_split_bwd(pre, grads, ctx.sizes)
        d_xn = dw = None
        if ctx.needs_input_grad[0]:
            d_xn = torch.empty((T, D), dtype=xn.dtype, device=xn.device)
            with _lib.timed("proj_gemm_uvqk_dgrad"):       # dx (T, Ntot) @ W^T: W is (N = D, K = Ntot) row-major
                _proj_gemm(dx, wc, T, D, Ntot, False, False, _lib.GEMM_EPI_PLAIN, d_xn)
        if ctx.needs_input_grad[1]:
            dw32 = _take_buf(ctx, "dw_buf", (D, Ntot), xn.device)
            with _lib.timed("proj_gemm_uvqk_wgrad"):       # xn^T (D, T) @ dx (T, Ntot), both read in place
                _proj_gemm(xn, dx, D, Ntot, T, True, True, _lib.GEMM_EPI_F32_ADD, dw32)
            dw = dw32 if ctx.w_dtype == torch.float32 else dw32.to(ctx.w_dtype)
        return (d_xn, dw, None, None) + (None,) * len(ctx.sizes)


def uvqk_projection(xn: torch.Tensor, w: torch.Tensor, sizes, w_cast: Optional[torch.Tensor] = None,
                    dw_buf: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, ...]:
    """``torch.split(F.silu(xn @ w), sizes, dim=1)`` (hstu.py:302-320), w (D, sum sizes) possibly an
    fp32 master of a bf16 xn.  ``w_cast``: w already in xn's dtype (``cast_many_bf16``); ``dw_buf``: a
    zero-filled fp32 tensor of w's shape the weight gradient is accumulated into (``zeros_many``)."""
    _lib.require_cuda(xn, w)
    sizes = [int(v) for v in sizes]
    if _gemm_ok(xn, xn.shape[1], w.shape[1]) and xn.shape[1] % 256 == 0 and all(v % 8 == 0 for v in sizes):
        return _UvqkProj.apply(xn, w, w_cast, dw_buf, *sizes)
    return silu_split(master_linear(xn, w, None, w_in_out=True), sizes)


class _OutProj(torch.autograd.Function):
    """F.linear(o_in, W, b) + residual (hstu.py:404-413) with bias and residual in the GEMM epilogue
    (one rounding to bf16 instead of two).  Backward: input gradient and fp32 weight gradient on the
    same kernel, bias gradient by a column-sum kernel; the residual's gradient is the incoming one."""

    @staticmethod
    def forward(ctx, o_in, w, b, res, w_cast, dw_buf, db_buf):
        o_in, res = _al16(o_in), _al16(res)
        T, Din = o_in.shape
        Dout = w.shape[0]
        if _usable_shadow(w, w_cast, o_in.dtype):
            wc = w_cast
        else:
            wc = (w if w.dtype == o_in.dtype else w.to(o_in.dtype)).contiguous()   # (N = Dout, K = Din) row-major
        ctx.dw_buf, ctx.db_buf = dw_buf, db_buf
        bf = None if b is None else (b if b.dtype == torch.float32 else b.float()).contiguous()
        out = torch.empty((T, Dout), dtype=o_in.dtype, device=o_in.device)
        with _lib.timed("proj_gemm_o_fwd"):
            _proj_gemm(o_in, wc, T, Dout, Din, False, False, _lib.GEMM_EPI_BIAS_RES, out, bias=bf, res=res)
        ctx.save_for_backward(o_in, wc)
        ctx.w_dtype = w.dtype
        ctx.b_dtype = None if b is None else b.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        o_in, wc = ctx.saved_tensors
        g = _al16(g)
        T, Din = o_in.shape
        Dout = wc.shape[0]
        d_in = dw = db = None
        if ctx.needs_input_grad[0]:
            d_in = torch.empty((T, Din), dtype=o_in.dtype, device=o_in.device)
            with _lib.timed("proj_gemm_o_dgrad"):          # g (T, Dout) @ W: W is (K = Dout, N = Din) row-major
                _proj_gemm(g, wc, T, Din, Dout, False, True, _lib.GEMM_EPI_PLAIN, d_in)
        if ctx.needs_input_grad[1]:
            dw32 = _take_buf(ctx, "dw_buf", (Dout, Din), g.device)
            with _lib.timed("proj_gemm_o_wgrad"):          # g^T (Dout, T) @ o_in (T, Din)
                _proj_gemm(g, o_in, Dout, Din, T, True, True, _lib.GEMM_EPI_F32_ADD, dw32)
            dw = dw32 if ctx.w_dtype == torch.float32 else dw32.to(ctx.w_dtype)
        if ctx.b_dtype is not None and ctx.needs_input_grad[2]:
            db32 = _take_buf(ctx, "db_buf", (Dout,), g.device)
            _lib.check(_lib.lib().grb_colsum_bf16(g.data_ptr(), g.stride(0), T, Dout, db32.data_ptr(),
                                                  _lib.stream_ptr(g.device)))
            db = db32 if ctx.b_dtype == torch.float32 else db32.to(ctx.b_dtype)
        return d_in, dw, db, (g if ctx.needs_input_grad[3] else None), None, None, None


def output_projection(o_in: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor],
                      residual: torch.Tensor, w_cast: Optional[torch.Tensor] = None,
                      dw_buf: Optional[torch.Tensor] = None, db_buf: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.linear(o_in, w, b) + residual`` (hstu.py:404-413), w / b possibly fp32 masters; ``w_cast`` /
    ``dw_buf`` / ``db_buf`` as in ``uvqk_projection``."""
    _lib.require_cuda(o_in, w, b, residual)
    if (_gemm_ok(o_in, o_in.shape[1], w.shape[0]) and w.shape[0] % 256 == 0 and w.shape[1] % 256 == 0
            and residual.dtype == o_in.dtype and residual.shape == (o_in.shape[0], w.shape[0])):
        return _OutProj.apply(o_in, w, b, residual, w_cast, dw_buf, db_buf)
    return master_linear(o_in, w, b) + residual


class _L2Norm(torch.autograd.Function):
    """x fp32 or bf16 -> y fp32 (the cast of the compute dtype is folded in); dx in x's dtype."""

    @staticmethod
    def forward(ctx, x, eps):
        shape = x.shape
        x2 = _rows_contiguous(x.reshape(-1, shape[-1]))
        rows, W = x2.shape
        y = torch.empty((rows, W), dtype=torch.float32, device=x.device)
        inv = torch.empty(rows, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().grb_l2norm_cast_fwd(x2.data_ptr(), _ld(x2), _lib.dtype_code(x2.dtype), y.data_ptr(), W,
                                                  inv.data_ptr(), rows, W, float(eps), _lib.stream_ptr(x.device)))
        ctx.save_for_backward(y, inv)
        ctx.shape = shape
        ctx.x_dtype = x2.dtype
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        rows, W = y.shape
        dy2 = dy.reshape(rows, W)
        if dy2.dtype != torch.float32:
            dy2 = dy2.float()
        dy2 = _rows_contiguous(dy2)
        dx = torch.empty((rows, W), dtype=ctx.x_dtype, device=y.device)
        _lib.check(_lib.lib().grb_l2norm_cast_bwd(y.data_ptr(), W, dy2.data_ptr(), _ld(dy2), inv.data_ptr(),
                                                  dx.data_ptr(), W, _lib.dtype_code(ctx.x_dtype), rows, W,
                                                  _lib.stream_ptr(y.device)))
        return dx.view(ctx.shape), None


def l2_normalize(x: torch.Tensor, eps: float) -> torch.Tensor:
    """x / clamp(||x||_2 over the last dim, min=eps): negative_sampler.py:31-37, postprocessors.py:47-55.
    One kernel forward, one backward (fp32 CUDA); other inputs take the reference's composite."""
    _lib.require_cuda(x)
    if x.dtype in (torch.float32, torch.bfloat16) and x.numel() > 0:
        return _L2Norm.apply(x, eps)       # bf16 rows (the encoder's compute dtype) come back as fp32
    return x / torch.clamp(torch.linalg.norm(x, ord=None, dim=-1, keepdim=True), min=eps)


class _TableGradState:
    """What the backward nodes of a TableGradScope share: the dense buffer and the one-element token.
    It references neither the proxy nor the table, so that no reference cycle keeps an autograd graph
    (and with it the table's gradient accumulator, pinned to the stream it was created on) alive after
    the step — a later CUDA-graph capture must not find one from an eager step on the default stream."""

    def __init__(self, shape, dtype, device) -> None:
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        self._buf = None
        self._tok = None

    def buffer(self) -> torch.Tensor:
        if self._buf is None:
            self._buf = torch.zeros(self.shape, dtype=torch.float32, device=self.device)
        return self._buf

    def token(self) -> torch.Tensor:
        if self._tok is None:
            self._tok = torch.zeros(1, dtype=self.dtype, device=self.device)
        return self._tok

    def take(self):
        buf, self._buf = self._buf, None
        return buf


class _TableSink(torch.autograd.Function):
    """See TableGradScope: stands between the table parameter and the readers of one step."""

    @staticmethod
    def forward(ctx, table, state):
        ctx.state = state
        return table.new_zeros(1)

    @staticmethod
    def backward(ctx, g):
        return ctx.state.take(), None


class TableGradScope:
    """One dense gradient buffer for ALL readers of an embedding table in a step.

    A train step reads the item table several times (encoder input, in-batch cache, supervision
    embeddings, local negatives).  With plain autograd every reader's backward zero-fills its own
    (V, D) fp32 tensor, scatters a few thousand rows into it, and the engine adds the dense tensors
    together: at C2 (131 263 x 256) that is 3 fills + 2 adds of 134 MB each, 0.16 ms of HBM traffic
    per step for nothing.  Readers given ``grad_scope=`` take a one-element proxy as their
    differentiable input instead of the table; their backward scatter-adds into the shared buffer
    (zeroed once, by whichever reader runs first) and hands the proxy a one-element token.  The engine
    runs ``_TableSink.backward`` after every reader that takes part in this backward pass has finished
    (its ordinary dependency counting), and that returns the buffer as the table's gradient."""

    def __init__(self, table: torch.Tensor):
        self.state = _TableGradState(table.shape, table.dtype, table.device)
        self.proxy = (_TableSink.apply(table, self.state)
                      if (table.requires_grad and torch.is_grad_enabled()) else None)


class _EmbeddingLookup(torch.autograd.Function):
    """weight[ids] with a scatter-add backward (one launch) instead of
    aten::embedding_dense_backward (sort + segmented reduce, ~30 launches)."""

    @staticmethod
    def forward(ctx, weight, ids, padding_idx, proxy=None, scope=None):
        _lib.require_cuda(weight, ids)
        ctx.save_for_backward(ids)
        ctx.shape = weight.shape
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        ctx.scope = scope.state if scope is not None else None
        return torch.embedding(weight, ids)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        V, D = ctx.shape
        g2 = g.reshape(-1, D)
        if g2.dtype != torch.float32 or g2.stride(-1) != 1:
            g2 = g2.float().contiguous()
        flat = ids.reshape(-1)
        if flat.dtype != torch.int64 or not flat.is_contiguous():
            flat = flat.to(torch.int64).contiguous()
        scope = ctx.scope
        dw = scope.buffer() if scope is not None else torch.zeros(V, D, device=g.device, dtype=torch.float32)
        _lib.check(_lib.lib().grb_rows_scatter_add(
            g2.data_ptr(), g2.stride(0), flat.data_ptr(), dw.data_ptr(), flat.numel(), D, V,
            ctx.padding_idx, _lib.stream_ptr(g.device)))
        if scope is not None:
            return None, None, None, scope.token(), None
        return dw, None, None, None, None


def embedding_lookup(weight: torch.Tensor, ids: torch.Tensor, padding_idx: Optional[int] = None,
                     grad_scope: Optional[TableGradScope] = None) -> torch.Tensor:
    """F.embedding(ids, weight, padding_idx) for fp32 CUDA tables (models/embeddings/
    embeddings.py:40-101); rows whose id equals padding_idx get no gradient.  ``grad_scope``: the
    table's gradient goes into the scope's shared buffer (TableGradScope)."""
    if weight.dtype != torch.float32:
        return torch.nn.functional.embedding(ids, weight, padding_idx)
    if grad_scope is not None and grad_scope.proxy is not None:
        return _EmbeddingLookup.apply(weight.detach(), ids, padding_idx, grad_scope.proxy, grad_scope)
    return _EmbeddingLookup.apply(weight, ids, padding_idx)


class _JaggedInput(torch.autograd.Function):
    """embeddings.py:94-97 + learnable_positional_embedding.py:42-58 + hstu.py:502 in one kernel each way."""

    @staticmethod
    def forward(ctx, table, pos, ids, offsets, rows, scale, p_drop, seed, out_dtype, proxy=None, scope=None):
        _lib.require_cuda(table, pos, ids, offsets)
        ctx.scope = scope.state if scope is not None else None
        if table.dtype != torch.float32 or pos.dtype != torch.float32:
            raise NotImplementedError("jagged_input: float32 tables only")
        table, pos = _rows_contiguous(table), _rows_contiguous(pos)
        ids, offsets = ids.contiguous(), offsets.contiguous()
        B, N = ids.shape
        D = table.shape[1]
        out = torch.empty((rows, D), dtype=out_dtype, device=table.device)
        a = _JaggedInput._args(table, pos, ids, offsets, rows, scale, p_drop, seed, out)
        _lib.check(_lib.lib().grb_jagged_input_fwd(C.byref(a), _lib.stream_ptr(table.device)))
        ctx.save_for_backward(ids, offsets, seed if seed is not None else ids.new_zeros(1))
        ctx.cfg = (tuple(table.shape), tuple(pos.shape), rows, scale, p_drop, seed is not None)
        return out

    @staticmethod
    def _args(table, pos, ids, offsets, rows, scale, p_drop, seed, io):
        a = _lib.JaggedInputArgs()
        a.B, a.N = ids.shape
        a.V, a.D = table.shape if table is not None else (0, io.shape[1])
        a.rows = rows
        a.dtype = _lib.dtype_code(io.dtype)
        a.index_bits = _lib.index_bits(offsets)
        if table is not None:
            a.table, a.ldt = table.data_ptr(), table.stride(0)
        if pos is not None:
            a.pos, a.ldp = pos.data_ptr(), pos.stride(0)
        a.ids, a.offsets = ids.data_ptr(), offsets.data_ptr()
        a.scale, a.p_drop = float(scale), float(p_drop)
        a.seed = _lib.ptr(seed)
        a.io, a.ldio = io.data_ptr(), io.stride(0)
        return a

    @staticmethod
    def backward(ctx, g):
        ids, offsets, seed = ctx.saved_tensors
        (V, D), (Np, _), rows, scale, p_drop, has_seed = ctx.cfg
        g = _rows_contiguous(g)
        scope = ctx.scope
        if scope is not None:
            d_table = scope.buffer() if ctx.needs_input_grad[9] else None
        else:
            d_table = torch.zeros((V, D), dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        d_pos = torch.zeros((Np, D), dtype=torch.float32, device=g.device) if ctx.needs_input_grad[1] else None
        a = _JaggedInput._args(None, None, ids, offsets, rows, scale, p_drop, seed if has_seed else None, g)
        a.V = V
        a.d_table, a.d_pos = _lib.ptr(d_table), _lib.ptr(d_pos)
        _lib.check(_lib.lib().grb_jagged_input_bwd(C.byref(a), _lib.stream_ptr(g.device)))
        if scope is not None:
            return (None, d_pos) + (None,) * 7 + (scope.token() if d_table is not None else None, None)
        return (d_table, d_pos) + (None,) * 9


def jagged_input(table: torch.Tensor, pos: torch.Tensor, ids: torch.Tensor, offsets: torch.Tensor,
                 rows: int, scale: float, p_drop: float = 0.0, seed: Optional[torch.Tensor] = None,
                 out_dtype: torch.dtype = torch.float32,
                 grad_scope: Optional[TableGradScope] = None) -> torch.Tensor:
    """Jagged encoder input (rows, D): ``dropout(table[ids] * scale + pos) * (ids != 0)`` for the valid
    positions only (embeddings.py:94-97, learnable_positional_embedding.py:42-58, hstu.py:502).
    ``rows`` >= offsets[-1]; extra rows are zero.  ``seed``: device int64 scalar when p_drop > 0."""
    if grad_scope is not None and grad_scope.proxy is not None:
        return _JaggedInput.apply(table.detach(), pos, ids, offsets, int(rows), float(scale), float(p_drop), seed,
                                  out_dtype, grad_scope.proxy, grad_scope)
    return _JaggedInput.apply(table, pos, ids, offsets, int(rows), float(scale), float(p_drop), seed, out_dtype)


def layer_norm_gate(x: torch.Tensor, gate: Optional[torch.Tensor], eps: float, p_drop: float = 0.0,
                    seed: Optional[torch.Tensor] = None, salt: int = 0) -> torch.Tensor:
    """gate * F.layer_norm(x, [W], eps=eps) without affine; gate=None gives the plain norm.
    p_drop > 0: dropout on the result (hstu.py:404-408), drawn inside the kernel from the device int64
    ``seed`` (one per step) and ``salt`` (one per call site) when ``ln_dropout_fusable(x)``, by F.dropout
    otherwise."""
    if p_drop > 0.0 and not (seed is not None and ln_dropout_fusable(x)
                             and (gate is None or ln_dropout_fusable(gate))):
        return torch.nn.functional.dropout(_LnGate.apply(x, gate, eps, 0.0, None, 0, False), p=p_drop, training=True)
    return _LnGate.apply(x, gate, eps, float(p_drop), seed if p_drop > 0.0 else None, salt, False)


def layer_norm_skip(x: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(F.layer_norm(x), x): the second output is x itself, routed through the same autograd node, for the
    layer's residual add (hstu.py:300 and :413) — see ``_LnGate``."""
    return _LnGate.apply(x, None, eps, 0.0, None, 0, True)


class _WeightedMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        out2 = torch.empty(2, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().grb_weighted_mean_fwd(x.data_ptr(), w.data_ptr(), x.numel(), out2.data_ptr(),
                                                    _lib.stream_ptr(x.device)))
        ctx.save_for_backward(w, out2)
        return out2[0]

    @staticmethod
    def backward(ctx, g):
        w, out2 = ctx.saved_tensors
        g = g.reshape(1).float().contiguous()
        gx = torch.empty(w.numel(), dtype=torch.float32, device=w.device)
        _lib.check(_lib.lib().grb_weighted_mean_bwd(w.data_ptr(), out2.data_ptr(), g.data_ptr(), w.numel(),
                                                    gx.data_ptr(), _lib.stream_ptr(w.device)))
        return gx, None


def weighted_mean(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``(x * w).sum() / w.sum()`` (autoregressive_losses.py:306) for 1-D float32 CUDA tensors, one launch
    each way; ``w`` (the supervision weights) receives no gradient.  Anything else takes the torch ops."""
    if (x.is_cuda and w.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and x.dim() == 1
            and x.shape == w.shape and x.is_contiguous() and w.is_contiguous() and not w.requires_grad):
        return _WeightedMean.apply(x, w)
    return (x * w).sum() / w.sum()


# --------------------------------------------------------------------------------------------
# fused sampled softmax  (negative_sampler.py:31-37,123-131,208-211; dot_product.py:61-64;
#                         autoregressive_losses.py:279-306)
# --------------------------------------------------------------------------------------------
class _SampledSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, p, table0, table1, idx0, idx1, pos_ids, neg_ids, l2_norm, eps, temperature,
                bf16_backward=False):
        _lib.require_cuda(q, p, table0, table1, idx0, idx1, pos_ids, neg_ids)
        for t in (q, p, table0, table1):
            if t is not None and t.dtype != torch.float32:
                raise NotImplementedError("sampled_softmax: float32 tensors only")
        q, p = _rows_contiguous(q), _rows_contiguous(p)
        table0 = _rows_contiguous(table0)
        table1 = _rows_contiguous(table1) if table1 is not None else None
        idx0 = idx0.contiguous()
        idx1 = idx1.contiguous() if idx1 is not None else None
        pos_ids, neg_ids = pos_ids.contiguous(), neg_ids.contiguous()
        n, D = q.shape
        R = idx0.shape[1]
        loss_rows = torch.empty(n, dtype=torch.float32, device=q.device)
        probs = torch.empty((n, R + 1), dtype=torch.float32, device=q.device)
        a = _SampledSoftmax._args(q, p, table0, table1, idx0, idx1, pos_ids, neg_ids, l2_norm, eps,
                                  temperature, probs)
        a.loss_rows = loss_rows.data_ptr()
        with _lib.timed("sampled_softmax_fwd"):
            _lib.check(_lib.lib().grb_sampled_softmax_fwd(C.byref(a), _lib.stream_ptr(q.device)))
        ctx.save_for_backward(q, p, table0, table1, idx0, idx1, pos_ids, neg_ids, probs)
        ctx.cfg = (l2_norm, eps, temperature)
        ctx.bf16_backward = bool(bf16_backward)
        return loss_rows

    @staticmethod
    def _args(q, p, table0, table1, idx0, idx1, pos_ids, neg_ids, l2_norm, eps, temperature, probs):
        a = _lib.SslArgs()
        a.n_rows, a.D = q.shape
        a.R = idx0.shape[1]
        a.d0 = table0.shape[1]
        a.d1 = table1.shape[1] if table1 is not None else 0
        a.l2_norm = 1 if l2_norm else 0
        a.dtype = _lib.GRB_F32
        a.l2_eps = float(eps)
        a.temperature = float(temperature)
        a.q, a.ldq_ = q.data_ptr(), _ld(q)
        a.p, a.ldp = p.data_ptr(), _ld(p)
        a.table0, a.ldt0 = table0.data_ptr(), _ld(table0)
        if table1 is not None:
            a.table1, a.ldt1 = table1.data_ptr(), _ld(table1)
            a.idx1 = idx1.data_ptr()
        a.idx0 = idx0.data_ptr()
        a.pos_ids, a.neg_ids = pos_ids.data_ptr(), neg_ids.data_ptr()
        a.probs = probs.data_ptr()
        return a

    @staticmethod
    def backward(ctx, g):
        q, p, table0, table1, idx0, idx1, pos_ids, neg_ids, probs = ctx.saved_tensors
        l2_norm, eps, temperature = ctx.cfg
        g = g.contiguous().float()
        dq = torch.empty(q.shape, dtype=torch.float32, device=q.device)
        dp = torch.empty(p.shape, dtype=torch.float32, device=q.device)
        n, D = q.shape
        R = idx0.shape[1]
        if (ctx.bf16_backward and table1 is None and not l2_norm and D % 8 == 0 and D <= 256 and n * R < 2 ** 31
                and os.environ.get("GRB_SSL_BWD_ATOMIC") != "1"):
            # one table of normalised rows (the in-batch cache): pairs counting-sorted by table row, both
            # halves of the backward read bf16 rows, no atomics on the table gradient, no zero fill
            dt0 = torch.empty(table0.shape, dtype=torch.float32, device=q.device)
            a = _SampledSoftmax._args(q, p, table0, None, idx0, None, pos_ids, neg_ids, l2_norm, eps,
                                      temperature, probs)
            a.g, a.dq, a.dp, a.dtable0 = g.data_ptr(), dq.data_ptr(), dp.data_ptr(), dt0.data_ptr()
            need = int(_lib.lib().grb_sampled_softmax_bwd_csr_workspace_bytes(n, R, D, table0.shape[0]))
            ws = torch.empty(need, dtype=torch.uint8, device=q.device)
            with _lib.timed("sampled_softmax_bwd"):
                _lib.check(_lib.lib().grb_sampled_softmax_bwd_csr(C.byref(a), table0.shape[0], ws.data_ptr(), need,
                                                                  _lib.stream_ptr(q.device)))
            return dq, dp, dt0, None, None, None, None, None, None, None, None, None
        dt0 = torch.zeros(table0.shape, dtype=torch.float32, device=q.device)
        dt1 = torch.zeros(table1.shape, dtype=torch.float32, device=q.device) if table1 is not None else None
        a = _SampledSoftmax._args(q, p, table0, table1, idx0, idx1, pos_ids, neg_ids, l2_norm, eps,
                                  temperature, probs)
        a.g, a.dq, a.dp = g.data_ptr(), dq.data_ptr(), dp.data_ptr()
        a.dtable0 = dt0.data_ptr()
        if dt1 is not None:
            a.dtable1 = dt1.data_ptr()
        with _lib.timed("sampled_softmax_bwd"):
            _lib.check(_lib.lib().grb_sampled_softmax_bwd(C.byref(a), _lib.stream_ptr(q.device)))
        return dq, dp, dt0, dt1, None, None, None, None, None, None, None, None


def sampled_softmax_rows(q: torch.Tensor, p: torch.Tensor, table0: torch.Tensor,
                         table1: Optional[torch.Tensor], idx0: torch.Tensor,
                         idx1: Optional[torch.Tensor], pos_ids: torch.Tensor,
                         neg_ids: torch.Tensor, l2_norm: bool, eps: float,
                         temperature: float, bf16_backward: bool = False) -> torch.Tensor:
    """Per-row sampled-softmax loss -log_softmax([q.p/T, masked q.e_r/T])[0], (N',) fp32.

    Negatives e_r = concat(table0[idx0[n, r]], table1[idx1[n, r]]) are gathered, optionally
    L2-normalised, dotted and reduced inside one kernel; (N', R, D) never exists.
    ``bf16_backward`` (for models whose activations are bf16 anyway): one table of normalised rows
    (the in-batch cache) takes the atomics-free backward of csrc/ssl_bwd_csr.cu, whose two sums read
    bf16 copies of q and of the table rows; the forward / the loss stay fp32."""
    return _SampledSoftmax.apply(q, p, table0, table1, idx0, idx1, pos_ids, neg_ids, l2_norm, eps,
                                 temperature, bf16_backward)


# --------------------------------------------------------------------------------------------
# fused MIPS top-k  (indexing/top_k.py:44-70)
# --------------------------------------------------------------------------------------------
_WS_CACHE: dict = {}


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


class MipsTopkCall:
    """One enqueued ``grb_mips_topk``: the outputs exist on the device, ordered on the stream; the
    overflow flag of the candidate workspace is copied to pinned host memory behind them.
    ``result()`` is where the host consumes the call: it waits for THIS call's flag only (not for the
    device), re-runs with the exact capacity when a row overflowed (rare: many equal scores) and
    returns the tensors.  A serving loop keeps one call in flight and asks for ``result()`` one batch
    later, where it reads the ids anyway; ``mips_topk`` is ``mips_topk_async(...).result()``."""

    def __init__(self, run, outs, flag, done):
        self._run, self._outs, self._flag, self._done = run, outs, flag, done

    def result(self):
        for _attempt in range(4):
            self._done.synchronize()
            overflow = int(self._flag[0])
            if overflow == 0:
                return self._outs
            # a row had more candidates than the workspace held: exact re-run
            self._outs, self._flag, self._done = self._run(overflow + 1024)
        raise RuntimeError("grb200 mips_topk: candidate workspace overflowed repeatedly")


def mips_topk_async(queries: torch.Tensor, items: torch.Tensor, item_ids: Optional[torch.Tensor],
                    k: int, invalid_ids: Optional[torch.Tensor] = None,
                    target_ids: Optional[torch.Tensor] = None) -> MipsTopkCall:
    """Enqueue the fused top-k (arguments as ``mips_topk``) without waiting for it."""
    _lib.require_cuda(queries, items, item_ids, invalid_ids, target_ids)
    if queries.dtype != items.dtype:
        raise ValueError("mips_topk: queries and items must share a dtype")
    queries, items = _rows_contiguous(queries), _rows_contiguous(items)
    B, D = queries.shape
    X = items.shape[0]
    if items.shape[1] != D:
        raise ValueError("mips_topk: embedding dims differ")
    if item_ids is not None:
        item_ids = item_ids.contiguous()
        if item_ids.dtype != torch.int64 or item_ids.numel() != X:
            raise ValueError("mips_topk: item_ids must be int64 with one id per item")
    if invalid_ids is not None:
        if invalid_ids.dim() != 2 or invalid_ids.shape[0] != B or invalid_ids.dtype != torch.int64:
            raise ValueError("mips_topk: invalid_ids must be (B, n) int64")
        invalid_ids = _rows_contiguous(invalid_ids)
        if invalid_ids.shape[1] == 0:
            invalid_ids = None
    if target_ids is not None:
        target_ids = target_ids.reshape(-1).contiguous()
        if target_ids.dtype != torch.int64 or target_ids.numel() != B:
            raise ValueError("mips_topk: target_ids must be int64 with one id per query")
    dev = queries.device
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    ranks = torch.empty(B, dtype=torch.int32, device=dev) if target_ids is not None else None
    outs = (out_s, out_i) if ranks is None else (out_s, out_i, ranks)
    a = _lib.MipsTopkArgs()
    a.B, a.X, a.D, a.k = B, X, D, k
    a.dtype = _lib.dtype_code(queries.dtype)
    a.queries, a.ldq = queries.data_ptr(), _ld(queries)
    a.items, a.ldi = items.data_ptr(), _ld(items)
    a.item_ids = _lib.ptr(item_ids)
    a.out_scores, a.out_ids = out_s.data_ptr(), out_i.data_ptr()
    if invalid_ids is not None:
        a.invalid_ids, a.ld_invalid, a.n_invalid = invalid_ids.data_ptr(), _ld(invalid_ids), invalid_ids.shape[1]
    if target_ids is not None:
        a.target_ids, a.out_ranks = target_ids.data_ptr(), ranks.data_ptr()
    keep = (queries, items, item_ids, invalid_ids, target_ids)   # alive until the re-run is ruled out

    def run(cap: int):
        status = torch.zeros(2, dtype=torch.int32, device=dev)
        a.status = status.data_ptr()
        a.sample_stride, a.cand_cap = 0, cap
        need = int(_lib.lib().grb_mips_topk_workspace_bytes(C.byref(a)))
        if need < 0:
            _lib.check(need)
        ws = _workspace(need, dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        with _lib.timed("mips_topk"):
            _lib.check(_lib.lib().grb_mips_topk(C.byref(a), _lib.stream_ptr(dev)))
        flag = _pinned_flag()
        flag.copy_(status, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        return outs, flag, done

    call = MipsTopkCall(run, *run(0))
    call._keep = keep
    return call


_FLAG_POOL: list = []


def _pinned_flag() -> torch.Tensor:
    """Pinned int32[2] landing pads for the overflow flags, recycled round-robin (a flag is read
    long before 64 further calls have been enqueued)."""
    if len(_FLAG_POOL) < 64:
        _FLAG_POOL.append(torch.zeros(2, dtype=torch.int32).pin_memory())
        return _FLAG_POOL[-1]
    f = _FLAG_POOL.pop(0)
    _FLAG_POOL.append(f)
    return f


def mips_topk(queries: torch.Tensor, items: torch.Tensor, item_ids: Optional[torch.Tensor],
              k: int, invalid_ids: Optional[torch.Tensor] = None,
              target_ids: Optional[torch.Tensor] = None):
    """Exact top-k of queries @ items.T, sorted descending, ties -> lowest item index.

    queries (B, D), items (X, D) row-major (same dtype, fp32 or bf16); item_ids (X,) int64 or
    None.  Returns (scores (B, k) fp32, ids (B, k) int64).

    invalid_ids (B, n) int64: ids that must not appear in row b's result — the filter of
    candidate_index.py:125-158 inside the selection kernel (thresholds for k' = k + n, the best k'
    sorted in shared memory, invalid ones dropped, first k written).  target_ids (B,) int64: a third
    output ranks (B,) int32 = 1 + position of the target in the row's result, k + 1 when absent
    (metrics/retrieval.py:45-55)."""
    return mips_topk_async(queries, items, item_ids, k, invalid_ids, target_ids).result()


class MipsTopkGraph:
    """Fixed-shape fused top-k replayed as ONE CUDA graph (serving: the same (B, D) query batch shape
    against the same corpus, call after call).  ``grb_mips_topk`` is a chain of ~11 small stream-ordered
    launches; at small batches (the HBM-bound regime) their launch latencies and the Python
    side of the call cost more than the kernels.  The graph owns static query / invalid-id / output
    buffers and its own workspace; ``__call__`` copies the inputs in, replays, and returns the static
    outputs (valid until the next call).  The overflow flag stays on the device: ``overflowed()`` reads
    it (host sync) — when set, fall back to ``mips_topk`` for that batch (exact re-run)."""

    def __init__(self, batch: int, items: torch.Tensor, item_ids: Optional[torch.Tensor], k: int,
                 n_invalid: int = 0, with_ranks: bool = False) -> None:
        _lib.require_cuda(items, item_ids)
        items = _rows_contiguous(items)
        dev = items.device
        X, D = items.shape
        self.items, self.item_ids = items, (item_ids.contiguous() if item_ids is not None else None)
        self.q = torch.zeros((batch, D), dtype=items.dtype, device=dev)
        self.invalid = torch.zeros((batch, n_invalid), dtype=torch.int64, device=dev) if n_invalid else None
        self.target = torch.zeros(batch, dtype=torch.int64, device=dev) if with_ranks else None
        self.scores = torch.empty((batch, k), dtype=torch.float32, device=dev)
        self.ids = torch.empty((batch, k), dtype=torch.int64, device=dev)
        self.ranks = torch.empty(batch, dtype=torch.int32, device=dev) if with_ranks else None
        self.status = torch.zeros(2, dtype=torch.int32, device=dev)
        a = _lib.MipsTopkArgs()
        a.B, a.X, a.D, a.k = batch, X, D, k
        a.dtype = _lib.dtype_code(items.dtype)
        a.queries, a.ldq = self.q.data_ptr(), _ld(self.q)
        a.items, a.ldi = items.data_ptr(), _ld(items)
        a.item_ids = _lib.ptr(self.item_ids)
        a.out_scores, a.out_ids = self.scores.data_ptr(), self.ids.data_ptr()
        a.status = self.status.data_ptr()
        if self.invalid is not None:
            a.invalid_ids, a.ld_invalid, a.n_invalid = self.invalid.data_ptr(), _ld(self.invalid), n_invalid
        if with_ranks:
            a.target_ids, a.out_ranks = self.target.data_ptr(), self.ranks.data_ptr()
        a.sample_stride, a.cand_cap = 0, 0
        need = int(_lib.lib().grb_mips_topk_workspace_bytes(C.byref(a)))
        if need < 0:
            _lib.check(need)
        self.ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        self._args = a
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):       # warm-up outside the capture (lazy module loading, attributes)
            _lib.check(_lib.lib().grb_mips_topk(C.byref(a), _lib.stream_ptr(dev)))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.status.zero_()
            _lib.check(_lib.lib().grb_mips_topk(C.byref(a), _lib.stream_ptr(dev)))

    def __call__(self, queries: torch.Tensor, invalid_ids: Optional[torch.Tensor] = None,
                 target_ids: Optional[torch.Tensor] = None):
        self.q.copy_(queries, non_blocking=True)
        if self.invalid is not None:
            self.invalid.copy_(invalid_ids, non_blocking=True)
        if self.target is not None:
            self.target.copy_(target_ids.reshape(-1), non_blocking=True)
        with _lib.timed("mips_topk"):
            self.graph.replay()
        return (self.scores, self.ids) if self.ranks is None else (self.scores, self.ids, self.ranks)

    def overflowed(self) -> bool:
        return int(self.status[0].item()) != 0


def topk_merge(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact top-k of per-row candidate lists (B, C): sorted descending, ties -> lowest id.
    This is the merge step after the all-gather of per-shard top-k."""
    _lib.require_cuda(cand_scores, cand_ids)
    cand_scores = cand_scores.contiguous().float()
    cand_ids = cand_ids.contiguous()
    B, cap = cand_scores.shape
    out_s = torch.empty((B, k), dtype=torch.float32, device=cand_scores.device)
    out_i = torch.empty((B, k), dtype=torch.int64, device=cand_scores.device)
    _lib.check(_lib.lib().grb_topk_select(
        cand_scores.data_ptr(), cand_ids.data_ptr(), None, B, cap, k, None, out_s.data_ptr(),
        out_i.data_ptr(), _lib.stream_ptr(cand_scores.device)))
    return out_s, out_i
