"""Generate tests/golden/*.pt by running the REAL reference modules.   *** TEST INFRASTRUCTURE ***

Run in the build container only (needs /root/reference; the GPU box has no reference tree):

    python oracle/make_golden.py

Imports the reference's hot-path modules from /root/reference/src (they run on CPU with their
own pure-PyTorch fallbacks — what ``trainer=cpu`` executes), drives them on small seeded inputs
and stores inputs, parameters (reference names) and outputs/gradients.  The fixtures pin
``oracle/reference_port.py`` (tests/test_oracle_golden.py) and, through it and directly, the
CUDA kernels (tests/test_*_gpu.py).
"""
from __future__ import annotations

import logging
import sys
from pathlib import Path

import torch

REF = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _import_reference():
    sys.path.insert(0, str(REF))
    logging.disable(logging.CRITICAL)  # the reference logs every fbgemm fallback
    from generative_recommenders_pl.models.utils import ops
    from generative_recommenders_pl.models.sequential_encoders import hstu
    from generative_recommenders_pl.models.indexing.top_k import MIPSBruteForceTopK
    from generative_recommenders_pl.models.indexing.candidate_index import CandidateIndex
    from generative_recommenders_pl.models.negatives_samples.negative_sampler import (
        InBatchNegativesSampler, LocalNegativesSampler)
    from generative_recommenders_pl.models.similarity.dot_product import DotProductSimilarity
    from generative_recommenders_pl.models.losses.autoregressive_losses import SampledSoftmaxLoss
    return dict(ops=ops, hstu=hstu, MIPSBruteForceTopK=MIPSBruteForceTopK,
                CandidateIndex=CandidateIndex, InBatch=InBatchNegativesSampler,
                Local=LocalNegativesSampler, Dot=DotProductSimilarity, SSL=SampledSoftmaxLoss)


def synth_timestamps(B, N, lengths, gen, big_gaps=False):
    """Unix-second style timestamps: increasing inside each history, 0 beyond it except index
    n_b which holds the 'target' timestamp (features.py:53-57)."""
    steps = torch.randint(1, 5000, (B, N), generator=gen)
    if big_gaps:
        steps = steps * torch.randint(1, 2000, (B, N), generator=gen)
    ts = 978_300_000 + torch.cumsum(steps, dim=1)
    out = torch.zeros(B, N, dtype=torch.int64)
    for b in range(B):
        n = int(lengths[b])
        out[b, :n] = ts[b, :n]
        if n < N:
            out[b, n] = ts[b, n - 1] + 100 if n > 0 else 0
    return out


def golden_ops(R):
    ops = R["ops"]
    g = {}
    # the reference's own test vectors, tests/test_ops.py:7-53
    l = torch.tensor([1, 2], dtype=torch.int32)
    g["kat_cumsum_in"], g["kat_cumsum_out"] = l, ops.asynchronous_complete_cumsum(l)
    x = torch.tensor([[1., 2., 3.], [4., 5., 6.]]).unsqueeze(-1)
    g["kat_d2j_in"], g["kat_d2j_out"] = x, ops.dense_to_jagged(x, g["kat_cumsum_out"])
    v = torch.tensor([1., 4., 5.]).unsqueeze(-1)
    g["kat_j2d_in"] = v
    g["kat_j2d_off"] = torch.tensor([0, 1, 3])
    g["kat_j2d_out"] = ops.jagged_to_padded_dense(v, g["kat_j2d_off"], 3, 0)
    # random ragged case, int64 offsets, with empty and full sequences
    gen = torch.Generator().manual_seed(1)
    lengths = torch.tensor([0, 7, 3, 12, 1, 0, 12, 5], dtype=torch.int64)
    dense = torch.randn(8, 12, 10, generator=gen)
    off = ops.asynchronous_complete_cumsum(lengths)
    jag = ops.dense_to_jagged(dense, off)
    g.update(r_lengths=lengths, r_dense=dense, r_offsets=off, r_jagged=jag,
             r_padded=ops.jagged_to_padded_dense(jag, off, 12, 0.0),
             r_padded_pad=ops.jagged_to_padded_dense(jag, off, 12, -1.5))
    lengths_nz = torch.tensor([4, 7, 3, 12, 1, 2, 12, 5], dtype=torch.int64)
    g["cur_lengths"] = lengths_nz
    g["cur_out"] = ops.get_current_embeddings(lengths_nz, dense)
    # mask_dense_by_aux_mask, tests/test_ops.py:56-139 (case 2: different lengths)
    d = torch.tensor([[[1, 1], [2, 2], [3, 3], [4, 4]], [[5, 5], [6, 6], [0, 0], [0, 0]]],
                     dtype=torch.float)
    m = torch.tensor([[False, True, False, True], [True, False, False, False]])
    ln = torch.tensor([4, 2])
    o, nl = ops.mask_dense_by_aux_mask(d, m, ln, 4)
    g.update(mask_dense=d, mask_aux=m, mask_lengths=ln, mask_out=o, mask_new_lengths=nl)
    return g


def golden_bias(R):
    hstu = R["hstu"]
    g = {}
    fn = lambda x: (torch.log(torch.abs(x).clamp(min=1)) / 0.301).long()  # noqa: E731  hstu.py:579
    # bucket(d) at every integer near each bucket edge and at large gaps
    edges = [int(math_ceil) for math_ceil in
             torch.ceil(torch.exp(0.301 * torch.arange(0, 129, dtype=torch.float64))).clamp(max=2 ** 62).tolist()]
    probe = sorted({max(0, e + d) for e in edges for d in range(-3, 4)} |
                   {0, 1, 2, 3, 10 ** 6, 10 ** 9, 2 ** 31, 2 ** 31 + 1, 2 ** 40, 2 ** 53 + 1, 2 ** 62})
    probe = torch.tensor(probe, dtype=torch.int64)
    g["bucket_probe"] = probe
    g["bucket_value"] = torch.clamp(fn(probe), 0, 128)
    gen = torch.Generator().manual_seed(2)
    B, N = 3, 20
    lengths = torch.tensor([20, 9, 1])
    ts = synth_timestamps(B, N, lengths, gen, big_gaps=True)
    torch.manual_seed(3)
    mod = hstu.RelativeBucketedTimeAndPositionBasedBias(max_seq_len=N, num_buckets=128,
                                                        bucketization_fn=fn)
    g.update(bias_ts=ts, bias_ts_w=mod._ts_w.detach().clone(),
             bias_pos_w=mod._pos_w.detach().clone(), bias_out=mod(ts).detach())
    return g


def _hstu_case(R, name, B, max_seq, out_len, D, H, dqk, dv, blocks, lengths, seed,
               normalization="rel_bias", linear_activation="silu", with_ts=True, **enc_kw):
    hstu = R["hstu"]
    N = max_seq + out_len
    torch.manual_seed(seed)
    enc = hstu.HSTU(max_sequence_len=max_seq, max_output_len=out_len, embedding_dim=D,
                    item_embedding_dim=D, num_blocks=blocks, num_heads=H, linear_dim=dv,
                    attention_dim=dqk, normalization=normalization, linear_config="uvqk",
                    linear_activation=linear_activation, linear_dropout_rate=0.2, attn_dropout_rate=0.0,
                    **enc_kw)
    enc.eval()
    gen = torch.Generator().manual_seed(seed + 1)
    lengths = torch.tensor(lengths, dtype=torch.int64)
    ts = synth_timestamps(B, N, lengths, gen)
    x = torch.randn(B, N, D, generator=gen)
    valid = (torch.arange(N).unsqueeze(0) < lengths.unsqueeze(1)).float().unsqueeze(-1)
    x = (x * valid).requires_grad_(True)
    y, _ = enc(past_lengths=lengths, user_embeddings=x, valid_mask=valid,
               past_payloads={"timestamps": ts} if with_ts else {})
    w = torch.randn(B, N, D, generator=gen)
    (y * w).sum().backward()
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items() if k != "_attn_mask"}
    grads = {k: p.grad.detach().clone() for k, p in enc.named_parameters()}
    return {f"{name}.cfg": torch.tensor([B, max_seq, out_len, D, H, dqk, dv, blocks]),
            f"{name}.lengths": lengths, f"{name}.ts": ts, f"{name}.x": x.detach().clone(),
            f"{name}.w": w, f"{name}.y": y.detach().clone(), f"{name}.dx": x.grad.detach().clone(),
            **{f"{name}.sd.{k}": v for k, v in sd.items()},
            **{f"{name}.grad.{k}": v for k, v in grads.items()}}


def golden_hstu(R):
    g = {}
    # tiny multi-head case with empty-ish / full / ragged sequences
    g.update(_hstu_case(R, "mh", B=5, max_seq=20, out_len=4, D=16, H=2, dqk=8, dv=8, blocks=2,
                        lengths=[1, 24, 7, 13, 2], seed=10))
    # ml-1m head shape (d = 50, one head), N = 211 as configs/model/hstu.yaml
    g.update(_hstu_case(R, "ml1m", B=3, max_seq=200, out_len=11, D=50, H=1, dqk=50, dv=50,
                        blocks=2, lengths=[200, 37, 129], seed=20))
    # tensor-core shaped heads (dqk = dv = 64), spans more than one 128-row tile
    g.update(_hstu_case(R, "h64", B=3, max_seq=150, out_len=11, D=128, H=2, dqk=64, dv=64,
                        blocks=1, lengths=[161, 130, 5], seed=30))
    return g


def golden_hstu_options(R):
    """The layer options of SequentialTransductionUnitJagged that no shipped config switches on
    (hstu.py:304-307 linear_activation="none", :398-402 concat_ua, :586-592 no relative bias)."""
    g = {}
    dims = dict(B=5, max_seq=20, out_len=4, D=16, H=2, dqk=8, dv=8, blocks=2, lengths=[1, 24, 7, 13, 2])
    g.update(_hstu_case(R, "ua", seed=90, concat_ua=True, **dims))
    g.update(_hstu_case(R, "noact", seed=91, linear_activation="none", **dims))
    # (without the bias module the reference only runs when no timestamps are passed: hstu.py:191-192)
    g.update(_hstu_case(R, "norab", seed=92, enable_relative_attention_bias=False, with_ts=False, **dims))
    g.update(_hstu_case(R, "ua64", B=3, max_seq=150, out_len=11, D=64, H=2, dqk=64, dv=64, blocks=1,
                        lengths=[161, 130, 5], seed=93, concat_ua=True))
    return g


def golden_hstu_softmax(R):
    """normalization="softmax_rel_bias" (hstu.py:337-384), forward and backward."""
    g = {}
    g.update(_hstu_case(R, "mh", B=5, max_seq=20, out_len=4, D=16, H=2, dqk=8, dv=8, blocks=2,
                        lengths=[1, 24, 7, 13, 2], seed=70, normalization="softmax_rel_bias"))
    g.update(_hstu_case(R, "h64", B=3, max_seq=60, out_len=4, D=64, H=2, dqk=64, dv=64, blocks=1,
                        lengths=[64, 30, 5], seed=80, normalization="softmax_rel_bias"))
    return g


def _hstu_incremental_case(R, name, B, max_seq, out_len, D, H, dqk, dv, blocks, lengths, seed,
                           keep_updated_cache=True):
    """Incremental path: a full pass with return_cache_states=True, then the LAST token of every
    sequence is replaced and recomputed through delta_x_offsets + cache (hstu.py:293-298,
    :151-177, :415-418).  The reference updates the caches in place, so copies go into the fixture."""
    hstu = R["hstu"]
    N = max_seq + out_len
    torch.manual_seed(seed)
    enc = hstu.HSTU(max_sequence_len=max_seq, max_output_len=out_len, embedding_dim=D,
                    item_embedding_dim=D, num_blocks=blocks, num_heads=H, linear_dim=dv,
                    attention_dim=dqk, normalization="rel_bias", linear_config="uvqk",
                    linear_activation="silu", linear_dropout_rate=0.2, attn_dropout_rate=0.0)
    enc.eval()
    gen = torch.Generator().manual_seed(seed + 1)
    lengths = torch.tensor(lengths, dtype=torch.int64)
    ts = synth_timestamps(B, N, lengths, gen)
    valid = (torch.arange(N).unsqueeze(0) < lengths.unsqueeze(1)).float().unsqueeze(-1)
    x = torch.randn(B, N, D, generator=gen) * valid
    kw = dict(past_lengths=lengths, valid_mask=valid, past_payloads={"timestamps": ts})
    with torch.no_grad():
        y0, cache = enc(user_embeddings=x, return_cache_states=True, **kw)
        cache0 = [tuple(t.detach().clone() for t in st) for st in cache]
        x2 = x.clone()
        x2[torch.arange(B), lengths - 1] = torch.randn(B, D, generator=gen)
        off = torch.zeros(B + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(lengths, 0)
        delta = (off[1:] - 1, lengths - 1)   # int64: index_copy_ in this torch rejects the int32 the docstring names
        y_inc, cache1 = enc(user_embeddings=x2, delta_x_offsets=delta, cache=cache,
                            return_cache_states=True, **kw)
        y_full, _ = enc(user_embeddings=x2, **kw)
    assert torch.allclose(y_inc, y_full, rtol=1e-4, atol=1e-5), (y_inc - y_full).abs().max()
    g = {f"{name}.cfg": torch.tensor([B, max_seq, out_len, D, H, dqk, dv, blocks]),
         f"{name}.lengths": lengths, f"{name}.ts": ts, f"{name}.x": x, f"{name}.x2": x2,
         f"{name}.delta0": delta[0], f"{name}.delta1": delta[1],
         f"{name}.y0": y0.clone(), f"{name}.y_inc": y_inc.clone(),
         **{f"{name}.sd.{k}": v.detach().clone() for k, v in enc.state_dict().items() if k != "_attn_mask"}}
    for i in range(blocks):
        for j, nm in enumerate(("v", "padded_q", "padded_k", "out")):
            g[f"{name}.cache0.{i}.{nm}"] = cache0[i][j]
            if keep_updated_cache:
                g[f"{name}.cache1.{i}.{nm}"] = cache1[i][j].detach().clone()
    return g


def golden_hstu_incremental(R):
    g = {}
    g.update(_hstu_incremental_case(R, "mh", B=5, max_seq=20, out_len=4, D=16, H=2, dqk=8, dv=8,
                                    blocks=2, lengths=[1, 24, 7, 13, 2], seed=50))
    g.update(_hstu_incremental_case(R, "h64", B=3, max_seq=150, out_len=11, D=64, H=2, dqk=64,
                                    dv=64, blocks=2, lengths=[161, 130, 5], seed=60,
                                    keep_updated_cache=False))     # (fixture size)
    return g


def golden_retrieval(R):
    g = {}
    gen = torch.Generator().manual_seed(40)
    X, D, B, k = 700, 24, 9, 10
    ids = torch.randperm(5000, generator=gen)[:X].sort().values + 1
    table = torch.nn.functional.normalize(torch.randn(X, D, generator=gen), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=-1)
    topk = R["MIPSBruteForceTopK"]()
    ci = R["CandidateIndex"](k=k, ids=ids, top_k_module=topk, embeddings=table.unsqueeze(0))
    s, i = topk(q, ci._embeddings_t, ci.ids, k=25, sorted=True)
    g.update(tk_ids=ids, tk_table=table, tk_q=q, tk_scores25=s, tk_ids25=i)
    # invalid ids: each row bans some of its own best hits plus padding zeros
    invalid = torch.zeros(B, 12, dtype=torch.int64)
    invalid[:, :6] = i[:, [0, 2, 3, 7, 11, 20]]
    oi, os_ = ci.get_top_k_outputs(q, invalid_ids=invalid)
    g.update(ci_invalid=invalid, ci_ids=oi, ci_scores=os_)

    # sampled softmax with the Local sampler over an nn.Embedding table
    n, R_, D2, V = 37, 16, 12, 60
    emb = torch.nn.Embedding(V + 1, D2, padding_idx=0)
    torch.manual_seed(41)
    emb.weight.data.normal_(0, 0.3)
    all_ids = list(range(1, V + 1))
    smp = R["Local"](l2_norm=True, l2_norm_eps=1e-6, all_item_ids=all_ids)
    smp._item_emb = emb
    sim, loss_fn = R["Dot"](), R["SSL"](num_to_sample=R_, softmax_temperature=0.05)
    out_emb = torch.nn.functional.normalize(torch.randn(n, D2, generator=gen), dim=-1).requires_grad_(True)
    sup_ids = torch.randint(1, V + 1, (n,), generator=gen)
    sup_emb = emb(sup_ids)
    sup_w = (torch.rand(n, generator=gen) > 0.2).float()
    torch.manual_seed(42)
    loss = loss_fn.jagged_forward(output_embeddings=out_emb, supervision_ids=sup_ids,
                                  supervision_embeddings=sup_emb, supervision_weights=sup_w,
                                  negatives_sampler=smp, similarity=sim)
    loss.backward()
    torch.manual_seed(42)  # replay the sampler's draw to record the ids it used
    neg_ids, _ = smp(sup_ids, R_)
    g.update(ssl_table=emb.weight.detach().clone(), ssl_out_emb=out_emb.detach().clone(),
             ssl_sup_ids=sup_ids, ssl_sup_w=sup_w, ssl_neg_ids=neg_ids, ssl_loss=loss.detach(),
             ssl_d_out_emb=out_emb.grad.clone(), ssl_d_table=emb.weight.grad.clone())

    # in-batch sampler with dedup: cache contents + loss for recorded offsets
    ib = R["InBatch"](l2_norm=True, l2_norm_eps=1e-6, dedup_embeddings=True)
    b_ids = torch.randint(0, 9, (30,), generator=gen)  # many duplicates, zeros = padding
    b_emb = emb(b_ids).detach()
    ib.process_batch(ids=b_ids, presences=(b_ids != 0), embeddings=b_emb)
    cid, cemb = ib.get_all_ids_and_embeddings()
    g.update(ib_ids=b_ids, ib_emb=b_emb, ib_cached_ids=cid, ib_cached_emb=cemb)
    return g


def main():
    R = _import_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])                    # e.g. `make_golden.py hstu_incremental`; default: all
    for name, fn in [("ops", golden_ops), ("bias", golden_bias), ("hstu", golden_hstu),
                     ("hstu_incremental", golden_hstu_incremental), ("hstu_softmax", golden_hstu_softmax), ("hstu_options", golden_hstu_options),
                     ("retrieval", golden_retrieval)]:
        if only and name not in only:
            continue
        data = fn(R)
        torch.save(data, OUT / f"{name}.pt")
        size = (OUT / f"{name}.pt").stat().st_size
        print(f"{name}: {len(data)} tensors, {size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
