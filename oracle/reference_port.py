"""CPU restatement of the reference's hot-path algorithms.   *** TEST INFRASTRUCTURE ***

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this file.  Nothing under ``mygenerativerecommenders_b200/``
imports it, and the product path never falls back to it.

Every function restates, in plain torch-on-CPU (dtype-generic: fp32 for baselines, fp64 for
tight parity), what the cited reference lines compute — the *reference's* formulation (padded
dense tensors, (B,H,N,N) scores, per-row Python loops for the jagged ops), not the fused
jagged formulation of the CUDA kernels.  Paths are relative to
/root/reference/src/generative_recommenders_pl/models/.

Parity pinning: ``tests/golden/*.pt`` hold inputs/outputs produced by importing the real
reference modules (``oracle/make_golden.py``, committed); ``tests/test_oracle_golden.py`` checks
this file against them, including the reference's own test vectors (tests/test_ops.py:7-139).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------
# a1-a3, a10: utils/ops.py
# ------------------------------------------------------------------------------------------
def complete_cumsum(lengths: torch.Tensor) -> torch.Tensor:
    """utils/ops.py:36-38 — [0, cumsum(lengths)], dtype of ``lengths``."""
    out = torch.zeros(lengths.numel() + 1, dtype=lengths.dtype)
    out[1:] = torch.cumsum(lengths, dim=0)
    return out


def dense_to_jagged(dense: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """utils/ops.py:60-64 — concatenate dense[b, :n_b] over b."""
    rows = []
    for b in range(offsets.numel() - 1):
        n_b = int(offsets[b + 1]) - int(offsets[b])
        rows.append(dense[b, :n_b])
    return torch.cat(rows, dim=0)


def jagged_to_padded_dense(values: torch.Tensor, offsets: torch.Tensor, max_len: int,
                           pad: float = 0.0) -> torch.Tensor:
    """utils/ops.py:100-114 — scatter jagged rows into a (B, max_len, ...) tensor of ``pad``.
    (The reference's fallback drops the dtype; the dtype is kept here so fp64 runs stay fp64.)"""
    B = offsets.numel() - 1
    out = torch.full((B, max_len, *values.shape[1:]), pad, dtype=values.dtype)
    for b in range(B):
        s, e = int(offsets[b]), int(offsets[b + 1])
        n_b = min(e - s, max_len)
        out[b, :n_b] = values[s:s + n_b]
    return out


def get_current_embeddings(lengths: torch.Tensor, enc: torch.Tensor) -> torch.Tensor:
    """utils/ops.py:183-187 — flattened gather of row lengths[b]-1."""
    B, N, D = enc.shape
    flat = (lengths - 1) + torch.arange(B, dtype=lengths.dtype) * N
    return enc.reshape(-1, D)[flat.long()]


def mask_dense_by_aux_mask(dense, aux_mask, lengths, max_len):
    """utils/ops.py:246-260."""
    off = complete_cumsum(lengths)
    jag = dense_to_jagged(dense, off)
    jm = dense_to_jagged(aux_mask, off)
    kept = jag[jm]
    new_len = aux_mask.int().sum(dim=1)
    return jagged_to_padded_dense(kept, complete_cumsum(new_len), max_len, 0.0), new_len


# ------------------------------------------------------------------------------------------
# a4: sequential_encoders/hstu.py:96-128
# ------------------------------------------------------------------------------------------
def bucketize_ts(diff: torch.Tensor, num_buckets: int = 128) -> torch.Tensor:
    """hstu.py:579-581 + :117-123: clamp(long(log(max(|d|,1)) / 0.301), 0, num_buckets).
    torch.log of an int64 tensor computes in float32, exactly like the reference."""
    return torch.clamp((torch.log(torch.abs(diff).clamp(min=1)) / 0.301).long(), 0, num_buckets)


def rel_bias(ts: torch.Tensor, ts_w: torch.Tensor, pos_w: torch.Tensor, N: int,
             num_buckets: int = 128) -> torch.Tensor:
    """hstu.py:106-128 — (B, N, N) bias: pos_w[N-1+j-i] + ts_w[bucket(ts[i+1]-ts[j])]."""
    B = ts.shape[0]
    i = torch.arange(N).view(N, 1)
    j = torch.arange(N).view(1, N)
    pos = pos_w[(N - 1) + j - i]                                   # == the pad/repeat trick
    ext = torch.cat([ts, ts[:, N - 1:N]], dim=1)                   # (B, N+1)
    buckets = bucketize_ts(ext[:, 1:].unsqueeze(2) - ext[:, :-1].unsqueeze(1), num_buckets)
    return pos.unsqueeze(0) + ts_w[buckets.view(-1)].view(B, N, N)


# ------------------------------------------------------------------------------------------
# a5: hstu.py:134-205
# ------------------------------------------------------------------------------------------
def hstu_attention(q, k, v, offsets, ts, ts_w, pos_w, N: int, H: int, dqk: int, dv: int):
    """Padded formulation of the reference: pad q,k,v -> einsum -> +bias -> SiLU/N -> causal
    mask -> einsum -> un-pad.  q,k (T, H*dqk), v (T, H*dv) -> (T, H*dv)."""
    B = offsets.numel() - 1
    pq = jagged_to_padded_dense(q, offsets, N).view(B, N, H, dqk)
    pk = jagged_to_padded_dense(k, offsets, N).view(B, N, H, dqk)
    pv = jagged_to_padded_dense(v, offsets, N).view(B, N, H, dv)
    s = torch.einsum("bnhd,bmhd->bhnm", pq, pk)
    if ts is not None:
        s = s + rel_bias(ts, ts_w, pos_w, N, ts_w.numel() - 1).unsqueeze(1).to(s.dtype)
    p = F.silu(s) / N
    causal = torch.tril(torch.ones(N, N, dtype=s.dtype))           # 1 - triu(ones, 1), :667
    p = p * causal
    out = torch.einsum("bhnm,bmhd->bnhd", p, pv).reshape(B, N, H * dv)
    return dense_to_jagged(out, offsets)


# ------------------------------------------------------------------------------------------
# a6-a8: hstu.py:266-423, :439-518, :633-672.  Parameters come as a state dict with the
# reference's names (``_hstu._attention_layers.{i}._uvqk`` ...).
# ------------------------------------------------------------------------------------------
def hstu_attention_softmax(q, k, v, offsets, bias, N: int, dqk: int):
    """normalization="softmax_rel_bias" (hstu.py:370-384): one attention over the full H*dqk width,
    softmax over ALL N columns of the padded row, causal mask applied after the softmax."""
    pq = jagged_to_padded_dense(q, offsets, N)
    pk = jagged_to_padded_dense(k, offsets, N)
    s = torch.einsum("bnd,bmd->bnm", pq, pk)
    if bias is not None:
        s = s + bias.to(s.dtype)
    p = F.softmax(s / math.sqrt(dqk), dim=-1) * torch.tril(torch.ones(N, N, dtype=s.dtype))
    return dense_to_jagged(torch.bmm(p, jagged_to_padded_dense(v, offsets, N)), offsets)


def stu_layer(x, offsets, ts, sd: Dict[str, torch.Tensor], prefix: str, N, H, dqk, dv,
              eps: float = 1e-6, dropout_p: float = 0.0, training: bool = False,
              linear_activation: str = "silu", concat_ua: bool = False,
              normalization: str = "rel_bias"):
    D = x.shape[1]
    xn = F.layer_norm(x, [D], eps=eps)
    mm = xn @ sd[prefix + "_uvqk"].to(x.dtype)
    if linear_activation == "silu":
        mm = F.silu(mm)
    u, v, q, k = torch.split(mm, [H * dv, H * dv, H * dqk, H * dqk], dim=1)
    has_bias = (prefix + "_rel_attn_bias._ts_w") in sd and ts is not None
    if normalization == "softmax_rel_bias":
        ts_w, pos_w = sd.get(prefix + "_rel_attn_bias._ts_w"), sd.get(prefix + "_rel_attn_bias._pos_w")
        bias = rel_bias(ts, ts_w, pos_w, N, ts_w.numel() - 1) if has_bias else None
        a = hstu_attention_softmax(q, k, v, offsets, bias, N, dqk)
    else:
        a = hstu_attention(q, k, v, offsets, ts if has_bias else None,
                           sd.get(prefix + "_rel_attn_bias._ts_w"),
                           sd.get(prefix + "_rel_attn_bias._pos_w"), N, H, dqk, dv)
    an = F.layer_norm(a, [H * dv], eps=eps)
    o_in = torch.cat([u, an, u * an], dim=-1) if concat_ua else u * an
    o_in = F.dropout(o_in, p=dropout_p, training=training)
    return F.linear(o_in, sd[prefix + "_o.weight"].to(x.dtype), sd[prefix + "_o.bias"].to(x.dtype)) + x


def stu_layer_cache_state(x, offsets, ts, sd, prefix, N, H, dqk, dv, eps: float = 1e-6):
    """The 4-tuple a layer returns with return_cache_states=True (hstu.py:420-423):
    (v jagged, padded_q (B,N,H*dqk), padded_k, layer output jagged).  Eval mode."""
    D = x.shape[1]
    mm = F.silu(F.layer_norm(x, [D], eps=eps) @ sd[prefix + "_uvqk"].to(x.dtype))
    _, v, q, k = torch.split(mm, [H * dv, H * dv, H * dqk, H * dqk], dim=1)
    out = stu_layer(x, offsets, ts, sd, prefix, N, H, dqk, dv, eps=eps)
    return (v.contiguous(), jagged_to_padded_dense(q, offsets, N, 0.0),
            jagged_to_padded_dense(k, offsets, N, 0.0), out)


def stu_layer_incremental(x, offsets, ts, sd, prefix, N, H, dqk, dv, delta, cache, eps: float = 1e-6):
    """The delta_x_offsets branch of SequentialTransductionUnitJagged.forward, as the reference
    computes it (hstu.py:293-298, :321-322, :151-177, :179-204, :397-401, :415-418): the new rows
    are written into the caches, the WHOLE padded attention is recomputed from them, and the rows
    delta[0] are kept.  Returns (layer output with the new rows, updated cache 4-tuple); the
    caches are cloned, not modified in place."""
    rows, pos = delta[0].long(), delta[1].long()
    cached_v, cached_q, cached_k, cached_out = (t.clone() for t in cache)
    B, D = offsets.numel() - 1, x.shape[1]
    xr = x[rows]
    mm = F.silu(F.layer_norm(xr, [D], eps=eps) @ sd[prefix + "_uvqk"].to(x.dtype))
    u, v, q, k = torch.split(mm, [H * dv, H * dv, H * dqk, H * dqk], dim=1)
    v_all = cached_v.index_copy_(0, rows, v)
    flat = pos + torch.arange(0, B * N, N)
    pq = cached_q.view(B * N, -1).index_copy_(0, flat, q).view(B, N, H, dqk)
    pk = cached_k.view(B * N, -1).index_copy_(0, flat, k).view(B, N, H, dqk)
    s = torch.einsum("bnhd,bmhd->bhnm", pq, pk)
    has_bias = (prefix + "_rel_attn_bias._ts_w") in sd and ts is not None
    if has_bias:
        ts_w, pos_w = sd[prefix + "_rel_attn_bias._ts_w"], sd[prefix + "_rel_attn_bias._pos_w"]
        s = s + rel_bias(ts, ts_w, pos_w, N, ts_w.numel() - 1).unsqueeze(1).to(s.dtype)
    p = F.silu(s) / N * torch.tril(torch.ones(N, N, dtype=s.dtype))
    pv = jagged_to_padded_dense(v_all, offsets, N).view(B, N, H, dv)
    a = dense_to_jagged(torch.einsum("bhnm,bmhd->bnhd", p, pv).reshape(B, N, H * dv), offsets)[rows]
    o_in = u * F.layer_norm(a, [H * dv], eps=eps)
    new = F.linear(o_in, sd[prefix + "_o.weight"].to(x.dtype), sd[prefix + "_o.bias"].to(x.dtype)) + xr
    out = cached_out.index_copy_(0, rows, new)
    return out, (v_all, pq.reshape(B, N, H * dqk), pk.reshape(B, N, H * dqk), out)


def hstu_forward_incremental(past_lengths, user_embeddings, timestamps, sd, num_blocks, H, dqk, dv,
                             delta, cache):
    """HSTU.forward with delta_x_offsets + cache (hstu.py:633-672 -> :439-478): every layer
    recomputes only the rows delta[0]; returns ((B, N, D), list of updated cache states)."""
    B, N, D = user_embeddings.shape
    off = complete_cumsum(past_lengths)
    x = dense_to_jagged(user_embeddings, off)
    states = []
    for i in range(num_blocks):
        x, st = stu_layer_incremental(x, off, timestamps, sd, f"_hstu._attention_layers.{i}.", N, H,
                                      dqk, dv, delta, cache[i])
        states.append(st)
    return jagged_to_padded_dense(x, off, N, 0.0), states


def hstu_cache_states(past_lengths, user_embeddings, timestamps, sd, num_blocks, H, dqk, dv):
    """The cache list HSTU.forward(return_cache_states=True) returns."""
    B, N, D = user_embeddings.shape
    off = complete_cumsum(past_lengths)
    x = dense_to_jagged(user_embeddings, off)
    states = []
    for i in range(num_blocks):
        st = stu_layer_cache_state(x, off, timestamps, sd, f"_hstu._attention_layers.{i}.", N, H, dqk, dv)
        states.append(st)
        x = st[3]
    return states


def hstu_forward(past_lengths, user_embeddings, timestamps, sd, num_blocks, H, dqk, dv,
                 dropout_p: float = 0.0, training: bool = False, **kw):
    """HSTU.forward (hstu.py:633-672): (B, N, D) -> (B, N, D)."""
    B, N, D = user_embeddings.shape
    off = complete_cumsum(past_lengths)
    x = dense_to_jagged(user_embeddings, off)
    for i in range(num_blocks):
        x = stu_layer(x, off, timestamps, sd, f"_hstu._attention_layers.{i}.", N, H, dqk, dv,
                      dropout_p=dropout_p, training=training, **kw)
    return jagged_to_padded_dense(x, off, N, 0.0)


# ------------------------------------------------------------------------------------------
# b1-b6: retrieval
# ------------------------------------------------------------------------------------------
def l2_normalize(x: torch.Tensor, eps: float) -> torch.Tensor:
    """negatives_samples/negative_sampler.py:33-36 == postprocessors/postprocessors.py:52-55."""
    return x / torch.clamp(torch.linalg.norm(x, dim=-1, keepdim=True), min=eps)


def mips_topk(queries: torch.Tensor, items: torch.Tensor, item_ids: Optional[torch.Tensor],
              k: int, chunk: int = 1 << 18) -> Tuple[torch.Tensor, torch.Tensor]:
    """indexing/top_k.py:62-70 — mm + topk + id gather, with the tie rule the north star fixes
    (lowest item index among equal scores; torch.topk's own tie order is arbitrary).  Scores are
    computed in fp32 from the stored values (bf16 inputs are up-cast), chunked over the corpus."""
    q = queries.float()
    B, X = q.shape[0], items.shape[0]
    best_s = torch.full((B, 0), 0.0)
    best_i = torch.zeros((B, 0), dtype=torch.int64)
    for lo in range(0, X, chunk):
        s = q @ items[lo:lo + chunk].float().t()
        idx = torch.arange(lo, lo + s.shape[1]).unsqueeze(0).expand(B, -1)
        cs = torch.cat([best_s, s], dim=1)
        ci = torch.cat([best_i, idx], dim=1)
        # order: score descending, index ascending  (stable sort on index-ordered input)
        order = torch.argsort(ci, dim=1, stable=True)
        cs, ci = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
        order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
    ids = best_i if item_ids is None else item_ids.reshape(-1)[best_i]
    return best_s, ids


def filter_invalid_topk(ids: torch.Tensor, scores: torch.Tensor, invalid_ids: torch.Tensor,
                        k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """indexing/candidate_index.py:142-158 — drop ids listed in invalid_ids[row], keep first k."""
    out_i, out_s = [], []
    for b in range(ids.shape[0]):
        banned = set(invalid_ids[b].tolist())
        keep = [c for c in range(ids.shape[1]) if int(ids[b, c]) not in banned][:k]
        out_i.append(ids[b, keep])
        out_s.append(scores[b, keep])
    return torch.stack(out_i), torch.stack(out_s)


def candidate_index_topk(queries, items, item_ids, k, invalid_ids=None):
    """CandidateIndex.get_top_k_outputs (candidate_index.py:107-164): returns (ids, scores)."""
    X = items.shape[0]
    n_inv = invalid_ids.shape[1] if invalid_ids is not None else 0
    k_prime = min(k + n_inv, X)
    s, i = mips_topk(queries, items, item_ids, k_prime)
    if invalid_ids is not None:
        i, s = filter_invalid_topk(i, s, invalid_ids, k)
    return i, s


def sampled_softmax_loss(out_emb, sup_ids, sup_emb, sup_w, neg_ids, neg_emb, temperature: float,
                         l2_eps: Optional[float] = 1e-6, neg_already_normalized: bool = False):
    """losses/autoregressive_losses.py:272-306 given the sampled ids and their *raw* gathered
    embeddings (N', R, D): normalise (negative_sampler.py:131), bmm (dot_product.py:61-64),
    /T, -5e4 collision mask, -log_softmax[:, 0], weighted mean.  Returns (loss, per-row loss)."""
    if l2_eps is not None:
        pos = l2_normalize(sup_emb, l2_eps)
        neg = neg_emb if neg_already_normalized else l2_normalize(neg_emb, l2_eps)
    else:
        pos, neg = sup_emb, neg_emb
    pos_logit = torch.bmm(pos.unsqueeze(1), out_emb.unsqueeze(2)).squeeze(2) / temperature
    neg_logit = torch.bmm(neg, out_emb.unsqueeze(2)).squeeze(2)
    neg_logit = torch.where(sup_ids.unsqueeze(1) == neg_ids,
                            torch.full_like(neg_logit, -5e4), neg_logit / temperature)
    rows = -F.log_softmax(torch.cat([pos_logit, neg_logit], dim=1), dim=1)[:, 0]
    return (rows * sup_w).sum() / sup_w.sum(), rows


def inbatch_process(ids, presences, embeddings, l2_eps: Optional[float], dedup: bool):
    """negatives_samples/negative_sampler.py:154-187 -> (cached_ids, cached_embeddings)."""
    valid_ids, valid_emb = ids[presences], embeddings[presences]
    if dedup:
        uniq, inv = torch.unique(valid_ids, sorted=False, return_inverse=True)
        rep = torch.empty(uniq.numel(), dtype=torch.int64)
        rep[inv] = torch.arange(valid_ids.numel())
        valid_ids, valid_emb = uniq, valid_emb[rep]
    return valid_ids, (l2_normalize(valid_emb, l2_eps) if l2_eps is not None else valid_emb)
