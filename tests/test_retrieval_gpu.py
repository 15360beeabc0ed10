"""b1-b6 on the GPU: fused MIPS top-k, candidate index, samplers and the fused sampled-softmax
loss against the reference's golden outputs and the CPU oracle.

Scores: fp32 tables agree to 1e-5 absolute on unit-norm embeddings; index sets must be identical
wherever adjacent score gaps exceed 2*tol, ties resolved to the lowest item index."""
import pytest
import torch

from mygenerativerecommenders_b200 import functional as GF
from mygenerativerecommenders_b200.candidate_index import CandidateIndex
from mygenerativerecommenders_b200.losses import SampledSoftmaxLoss
from mygenerativerecommenders_b200.negative_sampler import (InBatchNegativesSampler,
                                                             LocalNegativesSampler)
from mygenerativerecommenders_b200.similarity import DotProductSimilarity
from mygenerativerecommenders_b200.top_k import MIPSBruteForceTopK
from oracle import reference_port as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check_topk(got_s, got_i, ref_s, ref_i, tol):
    got_s, got_i = got_s.cpu().float(), got_i.cpu()
    assert torch.allclose(got_s, ref_s, atol=tol), (got_s - ref_s).abs().max()
    assert (got_s[:, 1:] <= got_s[:, :-1]).all(), "not sorted"
    # identical ids wherever the neighbouring gaps are > 2*tol; elsewhere same score is enough
    B, k = ref_i.shape
    for b in range(B):
        if torch.equal(got_i[b], ref_i[b]):
            continue
        for c in range(k):
            if got_i[b, c] != ref_i[b, c]:
                lo, hi = max(c - 1, 0), min(c + 1, k - 1)
                gap = min((ref_s[b, lo] - ref_s[b, c]).abs() if lo != c else 1e9,
                          (ref_s[b, c] - ref_s[b, hi]).abs() if hi != c else 1e9)
                assert gap <= 2 * tol, f"row {b} col {c}: ids differ across a gap of {gap}"


def test_golden_topk_and_candidate_index(golden):
    g = golden("retrieval")
    ids, table, q = g["tk_ids"].to(DEV), g["tk_table"].to(DEV), g["tk_q"].to(DEV)
    ci = CandidateIndex(k=10, ids=ids, top_k_module=MIPSBruteForceTopK(),
                        embeddings=table.unsqueeze(0)).to(DEV)
    assert ci._embeddings_t.shape == (table.shape[1], table.shape[0])
    assert ci.embeddings.shape == (1, *table.shape)
    s, i = ci._top_k_module(q, ci._embeddings_t, ci.ids, k=25, sorted=True)
    _check_topk(s, i, g["tk_scores25"], g["tk_ids25"], 1e-5)
    oi, os_ = ci.get_top_k_outputs(q, invalid_ids=g["ci_invalid"].to(DEV))
    assert torch.equal(oi.cpu(), g["ci_ids"])
    assert torch.allclose(os_.cpu(), g["ci_scores"], atol=1e-5)


def test_ties_resolve_to_lowest_index():
    base = torch.nn.functional.normalize(torch.randn(40, 16, generator=torch.Generator().manual_seed(3)), dim=-1)
    items = base.repeat(25, 1)                     # every item appears 25 times: massive ties
    q = base[:5].clone()
    s, i = GF.mips_topk(q.to(DEV), items.to(DEV), None, 60)
    rs, ri = O.mips_topk(q, items, None, 60)
    assert torch.equal(i.cpu(), ri)
    assert torch.allclose(s.cpu(), rs, atol=1e-6)


@pytest.mark.parametrize("B,X,D,k,dtype", [
    (128, 3883, 50, 411, torch.float32),           # C1: ml-1m corpus, k' = 200 + 211
    (128, 3883, 50, 411, torch.bfloat16),
    (17, 100_000, 64, 261, torch.float32),         # C3-shaped slice, k' = 200 + 61
    (130, 50_000, 256, 200, torch.bfloat16),       # C4-shaped slice: D = 256 bf16, 2 query tiles
    (3, 700, 24, 700, torch.float32),              # k == X
    (1, 129, 8, 1, torch.float32),
])
def test_random_corpora_vs_oracle(B, X, D, k, dtype):
    gen = torch.Generator().manual_seed(X + k)
    items = torch.nn.functional.normalize(torch.randn(X, D, generator=gen), dim=-1).to(dtype)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=-1).to(dtype)
    ids = torch.arange(1, X + 1) * 3
    s, i = GF.mips_topk(q.to(DEV), items.to(DEV), ids.to(DEV), k)
    rs, ri = O.mips_topk(q, items, ids, k)
    _check_topk(s, i, rs, ri, 1e-5 if dtype == torch.float32 else 2e-5)


def test_adversarial_order_forces_exact_rerun():
    # scores increase with the item index: the strided sample under-estimates tau badly,
    # candidate lists overflow, and the wrapper must still return the exact answer
    X, D, B, k = 60_000, 8, 4, 50
    items = torch.zeros(X, D)
    items[:, 0] = torch.linspace(-1, 1, X)
    q = torch.zeros(B, D)
    q[:, 0] = torch.tensor([1.0, 0.5, -1.0, 2.0])
    s, i = GF.mips_topk(q.to(DEV), items.to(DEV), None, k)
    rs, ri = O.mips_topk(q, items, None, k)
    assert torch.equal(i.cpu(), ri)


@pytest.mark.parametrize("B,X,D,k,n_inv,dtype", [
    (128, 3883, 50, 200, 211, torch.float32),      # C1/C2: k' = 411 inside the kernel, 200 written
    (128, 100_000, 64, 200, 61, torch.bfloat16),   # C3-shaped: past ids (128, 61)
    (9, 5_000, 32, 10, 3, torch.float32),
    (4, 900, 16, 100, 700, torch.float32),         # most of the best entries are invalid
])
def test_invalid_ids_filtered_inside_the_selection(B, X, D, k, n_inv, dtype):
    """f3: candidate_index.py:125-158 (over-select k + N, compare (B, k', N), keep the first k) against
    the fused selection that writes k entries; the target's rank against metrics/retrieval.py:45-55."""
    gen = torch.Generator().manual_seed(B * 7 + n_inv)
    items = torch.nn.functional.normalize(torch.randn(X, D, generator=gen), dim=-1).to(dtype)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=-1).to(dtype)
    ids = torch.randperm(4 * X, generator=gen)[:X] + 1
    ri_full, rs_full = O.candidate_index_topk(q, items, ids, min(k + n_inv, X))
    # invalid lists drawn mostly from each row's own best ids (so the filter bites), 0 = padding
    invalid = torch.zeros(B, n_inv, dtype=torch.int64)
    for b in range(B):
        pick = torch.randperm(ri_full.shape[1], generator=gen)[: n_inv // 2]
        invalid[b, : pick.numel()] = ri_full[b, pick]
        invalid[b, pick.numel(): n_inv - 1] = ids[torch.randint(0, X, (n_inv - 1 - pick.numel(),), generator=gen)]
    ri, rs = O.candidate_index_topk(q, items, ids, k, invalid_ids=invalid)
    target = ri[:, min(k - 1, 4)].clone()
    target[::3] = ids.max() + 7                    # absent -> rank k + 1
    target[1::3] = invalid[1::3, 0]                # filtered out -> absent as well
    s, i, ranks = GF.mips_topk(q.to(DEV), items.to(DEV), ids.to(DEV), k, invalid_ids=invalid.to(DEV),
                               target_ids=target.to(DEV))
    assert i.shape == (B, k)
    _check_topk(s, i, rs, ri, 1e-5 if dtype == torch.float32 else 2e-5)
    assert not (i.cpu().unsqueeze(2) == invalid.unsqueeze(1)).any()
    # metrics/retrieval.py:45-55 on the kernel's own ids
    _, rank_idx = torch.max(torch.cat([i.cpu(), target.unsqueeze(1)], dim=1) == target.unsqueeze(1), dim=1)
    assert torch.equal(ranks.cpu().long(), rank_idx + 1)
    from mygenerativerecommenders_b200.metrics import RetrievalMetrics
    m_ids, m_rk = RetrievalMetrics(k, [1, 10, min(k, 50)]), RetrievalMetrics(k, [1, 10, min(k, 50)])
    m_ids.update(i, target.to(DEV).unsqueeze(1))
    m_rk.update_ranks(ranks)
    assert all(torch.equal(v, m_rk.compute()[key]) for key, v in m_ids.compute().items())
    # the module boundary takes the fused path and returns (ids, scores)
    ci = CandidateIndex(k=k, ids=ids, top_k_module=MIPSBruteForceTopK(),
                        embeddings=items.to(DEV).unsqueeze(0)).to(DEV)
    oi, os_ = ci.get_top_k_outputs(q.to(DEV), invalid_ids=invalid.to(DEV))
    assert torch.equal(oi, i) and torch.equal(os_.float(), s)


def _launches(fn):
    from mygenerativerecommenders_b200 import _lib
    n0 = _lib.launch_count()
    out = fn()
    return out, _lib.launch_count() - n0


@pytest.mark.parametrize("B,X,D,k,n_inv", [
    (128, 700_000, 64, 200, 61),       # C3 (BASELINE.json configs[2]) at full size: k' = 261
    (37, 150_001, 128, 100, 0),        # ragged last item tile, rows 37..127 of the query block unused
    (128, 1_000_000, 256, 200, 0),     # C4's item width, group maxima per half tile (G = 64)
    (1, 200_000, 64, 10, 5),
    (96, 333_333, 192, 411, 211),      # C1/C2's k' = 200 + 211
])
def test_small_batch_plan_vs_oracle(B, X, D, k, n_inv):
    """One query block against a corpus that streams from HBM once (mips_small.cu: group-maxima sample,
    private sub-lists, register-resident select): exact ids against the oracle, in four launches."""
    gen = torch.Generator().manual_seed(B + X + k)
    items = torch.nn.functional.normalize(torch.randn(X, D, generator=gen), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=-1).to(torch.bfloat16)
    ids = torch.randperm(2 * X, generator=gen)[:X] + 1
    invalid = None
    if n_inv:
        ri_full, _ = O.candidate_index_topk(q, items, ids, k + n_inv)
        invalid = torch.zeros(B, n_inv, dtype=torch.int64)
        for b in range(B):       # half of each list from the row's own best ids, so that the filter bites
            pick = torch.randperm(ri_full.shape[1], generator=gen)[: n_inv // 2]
            invalid[b, : pick.numel()] = ri_full[b, pick]
    ri, rs = O.candidate_index_topk(q, items, ids, k, invalid_ids=invalid)
    target = ri[:, min(k - 1, 3)].clone()
    target[::2] = 2 * X + 9                          # absent -> rank k + 1
    qd, itd, idd = q.to(DEV), items.to(DEV), ids.to(DEV)
    invd = invalid.to(DEV) if invalid is not None else None
    GF.mips_topk(qd, itd, idd, k, invalid_ids=invd, target_ids=target.to(DEV))          # workspace warm-up
    (s, i, ranks), n = _launches(lambda: GF.mips_topk(qd, itd, idd, k, invalid_ids=invd,
                                                      target_ids=target.to(DEV)))
    assert n == 4, f"{n} launches: the small-batch plan was not taken"
    _check_topk(s, i, rs, ri, 2e-5)
    _, rank_idx = torch.max(torch.cat([i.cpu(), target.unsqueeze(1)], dim=1) == target.unsqueeze(1), dim=1)
    assert torch.equal(ranks.cpu().long(), rank_idx + 1)
    if invalid is not None:
        assert not (i.cpu().unsqueeze(2) == invalid.unsqueeze(1)).any()


def test_small_batch_plan_ties_and_overflow_fall_back_exactly():
    """Equal scores straddling the cut resolve to the lowest index inside the register-resident select;
    an item order (or a tie class) that overflows the private sub-lists sets the status word and the
    wrapper's re-run — the phased plan with an explicit capacity — still returns the exact answer."""
    gen = torch.Generator().manual_seed(5)
    X, D, B, k = 120_000, 64, 16, 150
    base = torch.nn.functional.normalize(torch.randn(6_000, D, generator=gen), dim=-1).to(torch.bfloat16)
    items = base.repeat(20, 1)                       # every item 20 times: k = 150 cuts through a tie class
    q = base[:B].clone()
    (s, i), n = _launches(lambda: GF.mips_topk(q.to(DEV), items.to(DEV), None, k))
    rs, ri = O.mips_topk(q, items, None, k)
    assert torch.equal(i.cpu(), ri) and torch.allclose(s.cpu(), rs, atol=2e-5)
    # scores increase with the index: the strided sample under-estimates tau, every sub-list of the last
    # CTAs overflows
    items = torch.zeros(X, D)
    items[:, 0] = torch.linspace(-1, 1, X)
    items = items.to(torch.bfloat16)
    q = torch.zeros(4, D)
    q[:, 0] = torch.tensor([1.0, 0.5, -1.0, 2.0])
    q = q.to(torch.bfloat16)
    (s, i), n = _launches(lambda: GF.mips_topk(q.to(DEV), items.to(DEV), None, 50))
    assert n > 4, "expected the overflow re-run"
    rs, ri = O.mips_topk(q, items, None, 50)
    assert torch.equal(i.cpu(), ri)


def test_async_call_defers_the_overflow_check_to_result():
    X, D, B, k = 60_000, 8, 4, 50
    items = torch.zeros(X, D)
    items[:, 0] = torch.linspace(-1, 1, X)         # adversarial order: the first attempt overflows
    q = torch.zeros(B, D)
    q[:, 0] = torch.tensor([1.0, 0.5, -1.0, 2.0])
    call = GF.mips_topk_async(q.to(DEV), items.to(DEV), None, k)
    s, i = call.result()
    assert torch.equal(i.cpu(), O.mips_topk(q, items, None, k)[1])


def test_topk_merge_matches_global_topk():
    gen = torch.Generator().manual_seed(8)
    B, G, kl, k = 33, 8, 200, 200
    s = torch.randn(B, G * kl, generator=gen)
    s[:, 100:140] = s[:, 300:340]                 # cross-shard ties
    ids = torch.randperm(10 ** 7, generator=gen)[:G * kl].unsqueeze(0).repeat(B, 1) + (1 << 33)
    ms, mi = GF.topk_merge(s.to(DEV), ids.to(DEV), k)
    order = torch.argsort(ids, dim=1, stable=True)
    cs, ci = torch.gather(s, 1, order), torch.gather(ids, 1, order)
    order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
    assert torch.equal(mi.cpu(), torch.gather(ci, 1, order))
    assert torch.equal(ms.cpu(), torch.gather(cs, 1, order))


def test_similarity_branches():
    sim = DotProductSimilarity()
    a = torch.randn(6, 8, device=DEV)
    out = sim(a, torch.randn(1, 10, 8, device=DEV), None, None)
    assert isinstance(out, tuple) and out[0].shape == (6, 10) and out[1] == {}
    assert sim(a, torch.randn(6, 10, 8, device=DEV), None, None).shape == (6, 10)
    assert sim(a, torch.randn(3, 10, 8, device=DEV), None, None).shape == (6, 10)


def _ssl_setup(g):
    table = g["ssl_table"].to(DEV)
    emb = torch.nn.Embedding(table.shape[0], table.shape[1], padding_idx=0).to(DEV)
    with torch.no_grad():
        emb.weight.copy_(table)
    smp = LocalNegativesSampler(l2_norm=True, l2_norm_eps=1e-6,
                                all_item_ids=list(range(1, table.shape[0]))).to(DEV)
    smp._item_emb = emb
    return emb, smp


def test_fused_sampled_softmax_vs_reference_golden(golden, monkeypatch):
    g = golden("retrieval")
    emb, smp = _ssl_setup(g)
    neg = g["ssl_neg_ids"].to(DEV)
    monkeypatch.setattr(smp, "_draw", lambda positive_ids, n: neg)   # the reference's CPU draw
    out_emb = g["ssl_out_emb"].to(DEV).requires_grad_(True)
    sup_ids = g["ssl_sup_ids"].to(DEV)
    loss_fn = SampledSoftmaxLoss(num_to_sample=neg.shape[1], softmax_temperature=0.05)
    loss = loss_fn.jagged_forward(output_embeddings=out_emb, supervision_ids=sup_ids,
                                  supervision_embeddings=emb(sup_ids),
                                  supervision_weights=g["ssl_sup_w"].to(DEV),
                                  negatives_sampler=smp, similarity=DotProductSimilarity())
    assert abs(loss.item() - g["ssl_loss"].item()) <= 1e-5 * abs(g["ssl_loss"].item())
    loss.backward()
    assert torch.allclose(out_emb.grad.cpu(), g["ssl_d_out_emb"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(emb.weight.grad.cpu(), g["ssl_d_table"], rtol=1e-4, atol=1e-6)
    # and the unfused module path (generic sampler.forward + similarity) agrees with the fused one
    emb.weight.grad = None
    out2 = g["ssl_out_emb"].to(DEV).requires_grad_(True)
    monkeypatch.setattr(smp, "fused_sample", lambda positive_ids, n: None)
    loss2 = loss_fn.jagged_forward(output_embeddings=out2, supervision_ids=sup_ids,
                                   supervision_embeddings=emb(sup_ids),
                                   supervision_weights=g["ssl_sup_w"].to(DEV),
                                   negatives_sampler=smp, similarity=DotProductSimilarity())
    assert abs(loss2.item() - loss.item()) <= 1e-5 * abs(loss.item())


def test_fused_two_table_concat_and_collisions():
    # this fork's embedding: concat(item_emb[id], year_emb[year_lookup[id]]) (embeddings.py:94-97)
    gen = torch.Generator().manual_seed(12)
    n, R, d0, d1, V = 300, 128, 25, 25, 40          # V small => many positive/negative collisions
    t0 = torch.randn(V + 1, d0, generator=gen) * 0.3
    t1 = torch.randn(12, d1, generator=gen) * 0.3
    years = torch.randint(0, 12, (V + 1,), generator=gen)
    q = torch.nn.functional.normalize(torch.randn(n, d0 + d1, generator=gen), dim=-1)
    pos_ids = torch.randint(1, V + 1, (n,), generator=gen)
    neg_ids = torch.randint(1, V + 1, (n, R), generator=gen)
    w = (torch.rand(n, generator=gen) > 0.1).float()
    assert (neg_ids == pos_ids.unsqueeze(1)).any()
    leaves = [x.clone().double().requires_grad_(True) for x in (q, t0, t1)]
    table = torch.cat([leaves[1][neg_ids], leaves[2][years[neg_ids]]], dim=-1)
    sup = torch.cat([leaves[1][pos_ids], leaves[2][years[pos_ids]]], dim=-1)
    ref, _ = O.sampled_softmax_loss(leaves[0], pos_ids, sup, w.double(), neg_ids, table, 0.05, 1e-6)
    ref.backward()
    gq, g0, g1 = (x.to(DEV).requires_grad_(True) for x in (q, t0, t1))
    sup_g = torch.cat([g0[pos_ids.to(DEV)], g1[years[pos_ids].to(DEV)]], dim=-1)
    rows = GF.sampled_softmax_rows(gq, O.l2_normalize(sup_g, 1e-6), g0, g1, neg_ids.to(DEV),
                                   years[neg_ids].to(DEV), pos_ids.to(DEV), neg_ids.to(DEV),
                                   True, 1e-6, 0.05)
    loss = (rows * w.to(DEV)).sum() / w.sum()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    loss.backward()
    for got, r, name in ((gq, leaves[0], "dq"), (g0, leaves[1], "dt0"), (g1, leaves[2], "dt1")):
        scale = r.grad.abs().max().item()
        assert (got.grad.cpu().double() - r.grad).abs().max().item() <= 2e-4 * scale, name


def test_inbatch_sampler_fused_vs_oracle(golden, monkeypatch):
    g = golden("retrieval")
    ib = InBatchNegativesSampler(l2_norm=True, l2_norm_eps=1e-6, dedup_embeddings=True)
    src = g["ib_emb"].to(DEV).requires_grad_(True)
    ids = g["ib_ids"].to(DEV)
    ib.process_batch(ids=ids, presences=(ids != 0), embeddings=src)
    cid, cemb = ib.get_all_ids_and_embeddings()
    ref = {int(i): e for i, e in zip(g["ib_cached_ids"], g["ib_cached_emb"])}
    assert sorted(ref) == sorted(cid.cpu().tolist())
    for i, e in zip(cid.cpu(), cemb.detach().cpu()):
        assert torch.allclose(e, ref[int(i)], atol=1e-6)
    gen = torch.Generator().manual_seed(5)
    n, R = 21, 16
    offs = torch.randint(0, cid.numel(), (n, R), generator=gen).to(DEV)
    monkeypatch.setattr(ib, "_draw", lambda positive_ids, k: offs)
    q = torch.nn.functional.normalize(torch.randn(n, cemb.shape[1], generator=gen), dim=-1).to(DEV).requires_grad_(True)
    sup_ids = cid[torch.randint(0, cid.numel(), (n,), generator=gen).to(DEV)]
    sup_emb = torch.randn(n, cemb.shape[1], generator=gen).to(DEV)
    w = torch.ones(n, device=DEV)
    loss = SampledSoftmaxLoss(R, 0.05).jagged_forward(q, sup_ids, sup_emb, w, ib, DotProductSimilarity())
    loss.backward()
    # oracle on CPU with the same offsets
    qc = q.detach().cpu().clone().requires_grad_(True)
    srcc = g["ib_emb"].clone().requires_grad_(True)
    c_ids, c_emb = O.inbatch_process(g["ib_ids"], g["ib_ids"] != 0, srcc, 1e-6, True)
    remap = {int(i): j for j, i in enumerate(c_ids)}
    o2 = torch.tensor([[remap[int(cid[o])] for o in row] for row in offs.cpu()])
    ref_loss, _ = O.sampled_softmax_loss(qc, sup_ids.cpu(), sup_emb.cpu(), w.cpu(), c_ids[o2],
                                         c_emb[o2], 0.05, 1e-6, neg_already_normalized=True)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert torch.allclose(q.grad.cpu(), qc.grad, rtol=1e-4, atol=1e-6)
    # gradient reaches the batch embeddings through the cache; per-id sums must agree
    # (which duplicate occurrence receives it is unspecified, negative_sampler.py:174-182)
    for item in set(g["ib_ids"].tolist()) - {0}:
        m = g["ib_ids"] == item
        assert torch.allclose(src.grad.cpu()[m].sum(0), srcc.grad[m].sum(0), rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("D,R,l2,bwd", [(256, 128, False, "csr"), (256, 128, False, "atomic"), (256, 40, True, ""),
                                         (128, 128, True, ""), (128, 33, False, "csr"), (128, 33, False, "atomic"),
                                         (512, 128, True, ""), (512, 20, False, ""), (64, 17, False, "csr")])
def test_fused_sampled_softmax_vector_path(D, R, l2, bwd, monkeypatch):
    """Single table with D = 128 / 256 / 512 takes the float4 kernels (32 negatives reduced together);
    same oracle as the generic path, incl. collisions and a ragged last batch of negatives.  A table of
    already-normalised rows (l2 False: the in-batch cache) takes the atomics-free backward
    (csrc/ssl_bwd_csr.cu: pairs counting-sorted by table row, bf16 gathers) when asked to (bf16_backward):
    its gradients carry the bf16 rounding of q and of the table rows inside the sums (4e-3 instead of 2e-4)."""
    if bwd == "atomic":          # the developer switch wins over the flag
        monkeypatch.setenv("GRB_SSL_BWD_ATOMIC", "1")
    gen = torch.Generator().manual_seed(D + R)
    n, V = 150, 60
    t0 = torch.randn(V + 1, D, generator=gen) * 0.3
    if not l2:   # the in-batch cache arrives normalised
        t0 = torch.nn.functional.normalize(t0, dim=-1)
    q = torch.nn.functional.normalize(torch.randn(n, D, generator=gen), dim=-1)
    pos_ids = torch.randint(1, V + 1, (n,), generator=gen)
    neg_ids = torch.randint(1, V + 1, (n, R), generator=gen)
    w = (torch.rand(n, generator=gen) > 0.1).float()
    assert (neg_ids == pos_ids.unsqueeze(1)).any()
    lq, lt = (x.clone().double().requires_grad_(True) for x in (q, t0))
    sup = torch.nn.functional.normalize(torch.randn(n, D, generator=gen), dim=-1)
    ref, _ = O.sampled_softmax_loss(lq, pos_ids, sup.double(), w.double(), neg_ids, lt[neg_ids], 0.05,
                                    1e-6, neg_already_normalized=not l2)
    ref.backward()
    gq, g0 = (x.to(DEV).requires_grad_(True) for x in (q, t0))
    rows = GF.sampled_softmax_rows(gq, sup.to(DEV), g0, None, neg_ids.to(DEV), None, pos_ids.to(DEV),
                                   neg_ids.to(DEV), l2, 1e-6, 0.05, bf16_backward=(bwd == "csr"))
    loss = (rows * w.to(DEV)).sum() / w.sum()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    loss.backward()
    tol = 4e-3 if bwd == "csr" else 2e-4
    for got, r, name in ((gq, lq, "dq"), (g0, lt, "dt0")):
        scale = r.grad.abs().max().item()
        assert (got.grad.cpu().double() - r.grad).abs().max().item() <= tol * scale, name
    assert torch.isfinite(g0.grad).all()


@pytest.mark.parametrize("direct", [False, True])
def test_inbatch_static_cache_matches_unique_cache(direct):
    """The sync-free in-batch cache (padded, count on the device) holds the same ids / embeddings as
    the torch.unique one, and its draws are uniform over [0, count)."""
    from mygenerativerecommenders_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, N, D = 11, 23, 16
    lengths = torch.randint(0, N - 1, (B,), generator=g)
    ids = torch.zeros(B, N, dtype=torch.int64)
    for b in range(B):
        ids[b, : lengths[b] + 1] = torch.randint(1, 40, (int(lengths[b]) + 1,), generator=g)
    ids = ids.to(DEV)
    table = torch.randn(41, D, generator=g).to(DEV)
    emb = table[ids]
    off = ops.asynchronous_complete_cumsum(lengths.to(DEV) + 1)
    total = int(lengths.sum()) + B
    dyn = InBatchNegativesSampler(True, 1e-6, True)
    sta = InBatchNegativesSampler(True, 1e-6, True)
    if direct:
        sta.max_item_id = 40            # direct-address build instead of the sort
    dyn.process_batch_prefix(ids, emb, off, total)
    sta.process_batch_prefix(ids, emb, off, total, static_shapes=True)
    c = int(sta._cached_count.item())
    assert c == dyn._cached_ids.numel() and sta._cached_ids.numel() == total
    # the same with rows padded to a bucket (what the CUDA-graph step does)
    pad = InBatchNegativesSampler(True, 1e-6, True)
    if direct:
        pad.max_item_id = 40
    pad.process_batch_prefix(ids, emb, off, total + 37, static_shapes=True, padded=True)
    p_ids, p_emb = pad.get_all_ids_and_embeddings()
    assert torch.equal(p_ids, dyn._cached_ids)
    assert torch.allclose(p_emb, dyn._cached_embeddings, atol=1e-7)
    a_ids, a_emb = sta.get_all_ids_and_embeddings()
    assert torch.equal(a_ids, dyn._cached_ids)
    assert torch.allclose(a_emb, dyn._cached_embeddings, atol=1e-7)
    pos = torch.zeros(4000, dtype=torch.int64, device=DEV)
    torch.manual_seed(77)
    d = sta._draw(pos, 128)
    assert d.min().item() >= 0 and d.max().item() == c - 1
    hist = torch.bincount(d.view(-1), minlength=c).float()
    expect = d.numel() / c
    assert (hist - expect).abs().max().item() < 6 * expect ** 0.5      # every bin within 6 sigma


def test_topk_graph_replay_equals_eager_call():
    """GF.MipsTopkGraph (one CUDA graph, static buffers) returns what the eager call returns, call after
    call with different queries; invalid ids and target ranks included."""
    gen = torch.Generator().manual_seed(11)
    X, D, B, k, n_inv = 60_000, 64, 40, 50, 7
    items = torch.nn.functional.normalize(torch.randn(X, D, generator=gen), dim=-1).to(torch.bfloat16).to(DEV)
    ids = (torch.randperm(2 * X, generator=gen)[:X] + 1).to(DEV)
    g = GF.MipsTopkGraph(B, items, ids, k, n_invalid=n_inv, with_ranks=True)
    for rep in range(3):
        q = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=-1).to(torch.bfloat16).to(DEV)
        inv = ids[torch.randint(0, X, (B, n_inv), generator=gen).to(DEV)]
        tgt = ids[torch.randint(0, X, (B,), generator=gen).to(DEV)]
        s, i, r = g(q, inv, tgt)
        es, ei, er = GF.mips_topk(q, items, ids, k, invalid_ids=inv, target_ids=tgt)
        assert not g.overflowed()
        assert torch.equal(i, ei) and torch.equal(s, es) and torch.equal(r, er)


def test_weighted_mean_one_launch_each_way():
    """autoregressive_losses.py:306, (loss * w).sum() / w.sum(): value and gradient against the torch ops."""
    from mygenerativerecommenders_b200 import _lib
    gen = torch.Generator().manual_seed(2)
    for n in (1, 33, 14_080, 300_001):
        x = torch.randn(n, generator=gen).to(DEV).requires_grad_(True)
        w = (torch.rand(n, generator=gen) > 0.3).float().to(DEV)
        w[0] = 1.0
        n0 = _lib.launch_count()
        y = GF.weighted_mean(x, w)
        (y * 3.0).backward()
        assert _lib.launch_count() - n0 == 2
        xr = x.detach().double().requires_grad_(True)
        yr = (xr * w.double()).sum() / w.double().sum()
        (yr * 3.0).backward()
        assert y.dim() == 0 and abs(y.item() - yr.item()) <= 1e-5 * max(1.0, abs(yr.item()))
        assert torch.allclose(x.grad.double(), xr.grad, rtol=1e-6, atol=1e-9)
    # anything the kernel does not take falls back to the torch expression
    xh = torch.randn(8, device=DEV, dtype=torch.bfloat16)
    assert GF.weighted_mean(xh, torch.ones(8, device=DEV, dtype=torch.bfloat16)).dtype == torch.bfloat16


@pytest.mark.parametrize("off_dtype", [torch.int64, torch.int32])
def test_inbatch_cache_from_the_table_is_the_sorted_distinct_id_list(off_dtype):
    """process_batch_table (grb_inbatch_distinct_ids: epoch-stamped membership table + one compaction) against
    torch.unique of the valid ids (negative_sampler.py:187-196), batch after batch on the same table — ids of
    an earlier batch must not reappear — with the caller's jagged offsets + rows_extra and a padded total."""
    from mygenerativerecommenders_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, N, D, V = 37, 61, 32, 5000
    table = torch.randn(V + 1, D, generator=g).to(DEV)
    table[0].zero_()
    smp = InBatchNegativesSampler(True, 1e-6, True)
    smp.max_item_id = V
    for rep in range(4):
        lengths = torch.randint(0, N - 1, (B,), generator=g)
        if rep == 2:
            lengths[:] = 0                                   # targets only
        ids = torch.zeros(B, N, dtype=torch.int64)
        for b in range(B):
            ids[b, : lengths[b] + 1] = torch.randint(1, V + 1 if rep % 2 else 200, (int(lengths[b]) + 1,), generator=g)
            ids[b, lengths[b] + 1:] = torch.randint(1, V + 1, (N - int(lengths[b]) - 1,), generator=g)   # junk past the prefix
        off = ops.asynchronous_complete_cumsum(lengths.to(DEV)).to(off_dtype)
        total = int(lengths.sum()) + B + (53 if rep % 2 else 0)          # an upper bound when padded
        assert smp.process_batch_table(ids.to(DEV), off, total, table, padded=bool(rep % 2), rows_extra=1)
        valid = torch.cat([ids[b, : lengths[b] + 1] for b in range(B)])
        want = torch.unique(valid)
        c = int(smp._cached_count.item())
        assert c == want.numel() and smp._cached_ids.numel() == total
        assert torch.equal(smp._cached_ids[:c].cpu(), want) and not smp._cached_ids[c:].any()
        ref = torch.nn.functional.normalize(table[want.to(DEV)], dim=-1, eps=1e-6)
        assert torch.allclose(smp._cached_embeddings[:c], ref, atol=1e-6) and not smp._cached_embeddings[c:].any()
