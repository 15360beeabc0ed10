"""Parity AT the benchmarked shapes (BASELINE.json configs), not at toy sizes:

* C2 (ml-20m-shaped): one whole train step at B=128, N=211, D=256, 4 heads x 64, 4 layers, bf16
  compute / fp32 master weights, the captured step graphs and the static-shape in-batch sampler
  that bench.py times -- loss and every parameter gradient against oracle/ref_step.py (the
  reference's formulation, fp32 CPU) with the same weights, batch and negative draws; then a
  3-step loss trajectory with FusedAdamW against the port stepping torch.optim.AdamW.
* C4 (retrieval-only): 4096 bf16 queries against a >= 1 M-item bf16 corpus, k=200: exact id
  equality against the chunked oracle (up-cast to fp32) wherever the score gap
  exceeds the fp32 accumulation noise; a corpus made of few distinct rows (thousands of exact
  ties per query) for the lowest-id rule.
* C5 (long sequence): one 8192-token sequence, H=2, forward + backward of the fused attention
  against the fp64 padded oracle.

Tolerances are the ones DESIGN.md §2 states: bf16 path ||d||_inf <= 1e-2 ||ref||_inf and
rel-L2 <= 5e-3 for outputs / the loss 1e-3; 2e-2 for gradients (both norms), except the relative-bias
tables ``_ts_w`` / ``_pos_w`` at 3e-2: each of their entries is a sum over ALL (sequence, head, i, j)
pairs of a bin, of signed terms that largely cancel, so the bf16 rounding noise of four bf16 layers
(1.3e-2 rel-L2 on the projection weights) shows up amplified there: 1.7-2.0e-2 rel-L2 and up to
2.4e-2 max with every kernel variant, the round-1 kernels and cuBLAS projections included
(benchmarks/probes/c2_parity_table.py).
"""
import pytest
import torch

from mygenerativerecommenders_b200 import functional as GF
from mygenerativerecommenders_b200 import hstu
from mygenerativerecommenders_b200.optim import FusedAdamW
from mygenerativerecommenders_b200.pipeline import (RetrievalConfig, RetrievalModel,
                                                    synthetic_batch, synthetic_item_ids)
from oracle import reference_port as O
from oracle.ref_step import RefRetrieval

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    inf = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)
    l2 = ((got - ref).norm() / max(ref.norm().item(), 1e-20)).item()
    return inf, l2


def c2_config(dropout=0.0):
    return RetrievalConfig(
        name="C2", num_items=131_262, max_sequence_length=200, gr_output_length=10,
        embedding_dim=256, num_blocks=4, num_heads=4, attention_dim=64, linear_dim=64,
        dropout=dropout, sampler="inbatch", num_negatives=128, temperature=0.05, top_k=200,
        split_year_embedding=False, compute_dtype=torch.bfloat16)


def _c2_pair(seed=42):
    cfg = c2_config()
    ids = synthetic_item_ids(26_744, cfg.num_items)
    torch.manual_seed(seed)
    m = RetrievalModel(cfg, ids)
    fp32 = RetrievalConfig(**{**cfg.__dict__, "compute_dtype": None})
    ref = RefRetrieval.from_state_dict(fp32, ids, m.state_dict()).train()
    return cfg, ids, m.to(DEV).train(), ref


def _inject_draws(m, raw):
    """The static in-batch sampler maps raw 62-bit draws onto [0, count) on the device; feed it
    fixed raw values so the oracle can replay the same picks."""
    smp = m.negatives_sampler

    def draw(positive_ids, n):
        count = smp._cached_count if smp._cached_count is not None else smp._cached_ids.size(0)
        return raw[: positive_ids.size(0)] % count
    smp._draw = draw


def _ref_draw(ref, row, picked_ids):
    """Translate the ids the GPU sampler picked into offsets of the oracle's in-batch pool."""
    lengths, pids, _ = ref.features(row)
    pids = pids.scatter(1, lengths.view(-1, 1), row["target_ids"].view(-1, 1))
    flat = pids.reshape(-1)
    cid, _ = O.inbatch_process(flat, flat != 0, ref.item_emb(flat), ref.cfg.l2_eps, True)
    order = torch.argsort(cid)
    pos = torch.searchsorted(cid[order], picked_ids)
    assert torch.equal(cid[order][pos], picked_ids)      # every pick is a member of the oracle's pool
    return order[pos]


def test_c2_train_step_graphs_loss_and_all_gradients_vs_oracle():
    cfg, ids, m, ref = _c2_pair()
    m.enable_step_graphs(row_granularity=1024)
    row = synthetic_batch(cfg, ids, 128, seed=1000)
    total = int(row["history_lengths"].sum())
    t_pad = -(-total // 1024) * 1024
    raw = torch.randint(0, 2 ** 40, (t_pad, cfg.num_negatives), device=DEV,
                        generator=torch.Generator(device=DEV).manual_seed(1))
    _inject_draws(m, raw)
    loss = m.training_loss({k: v.clone() for k, v in row.items()}, total_length=total)
    assert len(m._step_graphs) == 1
    loss.backward()
    smp = m.negatives_sampler
    picked = smp._cached_ids[(raw[:total] % smp._cached_count)].cpu()
    loss_ref = ref.training_loss(row, neg_draw=_ref_draw(ref, row, picked))
    loss_ref.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    ref_grads = {k.replace("|", "."): p.grad for k, p in ref.params.items()}
    checked = 0
    for k, p in m.named_parameters():
        r = ref_grads[k]
        if p.grad is None:
            assert r is None or r.abs().max() == 0, k
            continue
        inf, l2 = _rel(p.grad, r)
        tol = 3e-2 if "_rel_attn_bias" in k else 2e-2
        assert inf <= tol and l2 <= tol, f"{k}: max {inf:.3e} rel-l2 {l2:.3e}"
        checked += 1
    assert checked >= 4 * 5 + 2     # uvqk, o.weight, o.bias, ts_w, pos_w per layer + table + pos emb


def test_c2_three_step_loss_trajectory_with_fused_adamw_vs_oracle():
    cfg, ids, m, ref = _c2_pair(seed=7)
    # one row bucket for all three batches (totals are 14 080 +- 600): the sampler attributes read
    # back below are the static buffers of the graph captured last
    m.enable_step_graphs(row_granularity=4096)
    opt = FusedAdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
    # the draw closure is baked into the captured graph: one buffer, refreshed in place every step
    raw = torch.empty((26 * 1024, cfg.num_negatives), dtype=torch.int64, device=DEV)
    _inject_draws(m, raw)
    smp = m.negatives_sampler
    got, want = [], []
    for step in range(3):
        row = synthetic_batch(cfg, ids, 128, seed=2000 + step)
        total = int(row["history_lengths"].sum())
        raw.copy_(torch.randint(0, 2 ** 40, raw.shape, device=DEV))
        loss = m.training_loss({k: v.clone() for k, v in row.items()}, total_length=total)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        picked = smp._cached_ids[(raw[:total] % smp._cached_count)].cpu()
        opt.step()
        loss_ref = ref.training_loss(row, neg_draw=_ref_draw(ref, row, picked))
        opt_ref.zero_grad(set_to_none=True)
        loss_ref.backward()
        opt_ref.step()
        got.append(loss.item())
        want.append(loss_ref.item())
    assert len(m._step_graphs) == 1
    for g, w in zip(got, want):
        assert abs(g - w) <= 2e-3 * abs(w), (got, want)
    assert want[2] < want[0]       # and it trains


# ---------------------------------------------------------------------------------------------
# C4: retrieval-only, full query batch against a >= 1 M corpus
# ---------------------------------------------------------------------------------------------
def _oracle_topk_chunked(q, items, k, chunk=1 << 17):
    """fp32 mm + exact (score desc, index asc) selection, chunked over the corpus."""
    qf = q.float()
    best_s = torch.full((q.shape[0], 0), 0.0)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64)
    for lo in range(0, items.shape[0], chunk):
        s = qf @ items[lo:lo + chunk].float().t()
        kk = min(s.shape[1], k + 64)
        ts, ti = torch.topk(s, kk, dim=1)
        best_s = torch.cat([best_s, ts], 1)
        best_i = torch.cat([best_i, ti + lo], 1)
        if best_s.shape[1] > 4 * (k + 64):
            best_s, sel = torch.topk(best_s, k + 64, dim=1)
            best_i = torch.gather(best_i, 1, sel)
    # final exact order: score descending, index ascending
    key = torch.argsort(best_i, dim=1, stable=True)
    best_s, best_i = torch.gather(best_s, 1, key), torch.gather(best_i, 1, key)
    order = torch.argsort(best_s, dim=1, descending=True, stable=True)
    return torch.gather(best_s, 1, order)[:, :k + 1], torch.gather(best_i, 1, order)[:, :k + 1]


def test_c4_full_query_batch_million_item_corpus_exact_ids():
    B, X, D, k = 4096, 1_000_000, 256, 200
    g = torch.Generator().manual_seed(0)
    items = torch.nn.functional.normalize(torch.randn(X, D, generator=g), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1).to(torch.bfloat16)
    item_ids = torch.arange(1, X + 1, dtype=torch.int64)
    s, i = GF.mips_topk(q.to(DEV), items.to(DEV), item_ids.to(DEV), k)
    s, i = s.cpu(), i.cpu()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    rs, ri = _oracle_topk_chunked(q, items, k)
    assert torch.allclose(s, rs[:, :k], atol=2e-6, rtol=0)          # fp32 accumulation order only
    # ids must agree wherever the neighbouring scores are further apart than that noise
    tol = 4e-6
    gap_prev = torch.cat([torch.full((B, 1), 1.0), rs[:, :k - 1] - rs[:, 1:k]], 1)
    gap_next = rs[:, :k] - rs[:, 1:k + 1]
    decided = (gap_prev > tol) & (gap_next > tol)
    assert decided.float().mean().item() > 0.9       # (measured 0.95: neighbour gaps below tol are rare but not negligible)
    assert torch.equal(i[decided], ri[:, :k][decided] + 1)
    # and as sets, per row, away from the k-th boundary
    assert (torch.sort(i, 1).values == torch.sort(ri[:, :k] + 1, 1).values).float().mean().item() > 0.9995


def test_c4_scale_many_exact_ties_lowest_id_wins():
    B, X, D, k = 4096, 1_048_576, 256, 200
    g = torch.Generator().manual_seed(1)
    proto = torch.nn.functional.normalize(torch.randn(512, D, generator=g), dim=-1).to(torch.bfloat16)
    pick = torch.randint(0, 512, (X,), generator=g)
    items = proto[pick]                                   # every row has ~2048 exact duplicates
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1).to(torch.bfloat16)
    s, i = GF.mips_topk(q.to(DEV), items.to(DEV), None, k)
    s, i = s.cpu(), i.cpu()
    # oracle: score every prototype, the best prototype of a query has > k copies, so the answer is
    # the k lowest indices among its copies
    ps = q.float() @ proto.float().t()                    # (B, 512)
    best = ps.argmax(1)
    counts = torch.bincount(pick, minlength=512)
    assert counts.min() > k
    order = torch.argsort(pick, stable=True)              # indices grouped by prototype, ascending
    starts = torch.cumsum(counts, 0) - counts
    want = torch.stack([order[starts[b]: starts[b] + k] for b in best.tolist()])
    assert torch.equal(i, want)
    assert torch.allclose(s, ps.max(1).values.unsqueeze(1).expand(B, k), atol=2e-6, rtol=0)


# ---------------------------------------------------------------------------------------------
# C5: one 8192-token sequence
# ---------------------------------------------------------------------------------------------
def test_c5_one_8192_token_sequence_forward_backward_vs_fp64_oracle():
    N, H, d = 8192, 2, 64
    gen = torch.Generator().manual_seed(5)
    lengths = torch.tensor([N])
    off = O.complete_cumsum(lengths)
    q, k, v = ((torch.randn(N, H * d, generator=gen) * 0.5).to(torch.bfloat16).float() for _ in range(3))
    ts = (978_300_000 + torch.cumsum(torch.randint(1, 5000, (1, N), generator=gen), dim=1))
    ts_w = torch.randn(129, generator=gen) * 0.5
    pos_w = torch.randn(2 * N - 1, generator=gen) * 0.5
    w = torch.randn(N, H * d, generator=gen).to(torch.bfloat16).float()
    thr = hstu.tabulate_bucket_thresholds(hstu._default_bucketization, 128).to(DEV)
    leaves = [t.to(DEV).to(torch.bfloat16).requires_grad_(True) for t in (q, k, v)]
    tw, pw = ts_w.to(DEV).requires_grad_(True), pos_w.to(DEV).requires_grad_(True)
    offd, tsd = off.to(DEV), ts.to(DEV)
    cache = GF.hstu_bucket_cache(offd, tsd, thr, N)
    out = GF.hstu_attention(leaves[0], leaves[1], leaves[2], offd, tsd, tw, pw, thr, N, H, d, d,
                            bucket_cache=cache)
    out.backward(w.to(DEV).to(torch.bfloat16))
    rl = [t.double().requires_grad_(True) for t in (q, k, v, ts_w, pos_w)]
    ref = O.hstu_attention(rl[0], rl[1], rl[2], off, ts, rl[3], rl[4], N, H, d, d)
    (ref * w.double()).sum().backward()
    inf, l2 = _rel(out, ref)
    assert inf <= 1e-2 and l2 <= 5e-3, ("fwd", inf, l2)
    for name, got, r in zip(("dq", "dk", "dv", "d_ts_w", "d_pos_w"), leaves + [tw, pw], rl):
        inf, l2 = _rel(got.grad, r.grad)
        assert inf <= 2e-2 and l2 <= 2e-2, (name, inf, l2)
