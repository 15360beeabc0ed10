"""End-to-end: the whole train / eval step through this package's modules on the GPU against the
CPU restatement of the reference's step (oracle/ref_step.py) with identical weights, batch and
pre-drawn negatives.  Covers a9 (caller-side jagged packing), both samplers and b3."""
import pytest
import torch

from mygenerativerecommenders_b200.pipeline import (RetrievalConfig, RetrievalModel,
                                                    synthetic_batch, synthetic_item_ids)
from oracle.ref_step import RefRetrieval

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _models(sampler, split_year, D=32, H=2):
    cfg = RetrievalConfig(name="t", num_items=500, max_sequence_length=40, gr_output_length=5,
                          embedding_dim=D, num_blocks=2, num_heads=H, attention_dim=D // H,
                          linear_dim=D // H, dropout=0.2, sampler=sampler, num_negatives=16,
                          top_k=20, split_year_embedding=split_year)
    ids = synthetic_item_ids(300, cfg.num_items, seed=1)
    torch.manual_seed(0)
    m = RetrievalModel(cfg, ids)
    if split_year:
        m.embeddings.year_lookup_table.copy_(torch.randint(0, 30, (cfg.num_items + 1,)))
    with torch.no_grad():  # make the table non-tiny so scores are well separated
        for p in m.embeddings.parameters():
            p.mul_(10.0)
    ref = RefRetrieval.from_state_dict(cfg, ids, m.state_dict()).eval()
    return cfg, ids, m.to(DEV).eval(), ref


@pytest.mark.parametrize("sampler,split_year", [("local", True), ("local", False), ("inbatch", False)])
def test_training_loss_and_gradients(sampler, split_year, monkeypatch):
    cfg, ids, m, ref = _models(sampler, split_year)
    row = synthetic_batch(cfg, ids, 6, seed=3, min_len=2)
    n_rows = int(row["history_lengths"].sum())
    gen = torch.Generator().manual_seed(4)
    if sampler == "local":
        draw = torch.randint(0, ids.numel(), (n_rows, cfg.num_negatives), generator=gen)
        monkeypatch.setattr(m.negatives_sampler, "_draw",
                            lambda p, n: m.negatives_sampler._all_item_ids[draw.to(DEV)])
        ref_draw = draw
    else:
        # offsets index the de-duplicated in-batch pool whose order differs between devices:
        # draw ids instead and translate per side
        lengths, pids, _ = ref.features(row)
        pids = pids.scatter(1, lengths.view(-1, 1), row["target_ids"].view(-1, 1))
        pool = torch.unique(pids[pids != 0])
        pick = pool[torch.randint(0, pool.numel(), (n_rows, cfg.num_negatives), generator=gen)]

        def gpu_draw(p, n):
            cid = m.negatives_sampler._cached_ids
            order = torch.argsort(cid)
            return order[torch.searchsorted(cid[order], pick.to(DEV))]
        monkeypatch.setattr(m.negatives_sampler, "_draw", gpu_draw)
        flat = pids.reshape(-1)
        from oracle import reference_port as O
        cid, _ = O.inbatch_process(flat, flat != 0, ref.item_emb(flat), cfg.l2_eps, True)
        order = torch.argsort(cid)
        ref_draw = order[torch.searchsorted(cid[order], pick)]
    loss = m.training_loss({k: v.clone() for k, v in row.items()}, total_length=n_rows)
    loss_ref = ref.training_loss(row, neg_draw=ref_draw)
    assert abs(loss.item() - loss_ref.item()) <= 2e-5 * abs(loss_ref.item())
    loss.backward()
    loss_ref.backward()
    ref_grads = {k.replace("|", "."): p.grad for k, p in ref.params.items()}
    checked = 0
    for k, p in m.named_parameters():
        if p.grad is None:
            assert ref_grads[k] is None or ref_grads[k].abs().max() == 0, k
            continue
        r = ref_grads[k]
        scale = max(r.abs().max().item(), 1e-8)
        assert (p.grad.cpu() - r).abs().max().item() <= 1e-3 * scale, k
        checked += 1
    assert checked >= 8


def test_retrieve_matches_reference_step():
    cfg, ids, m, ref = _models("local", True)
    row = synthetic_batch(cfg, ids, 9, seed=5, min_len=2)
    got_ids, got_scores = m.retrieve(row)
    ref_ids, ref_scores = ref.retrieve(row)
    assert got_ids.shape == (9, cfg.top_k)
    assert torch.allclose(got_scores.cpu(), ref_scores, atol=2e-5)
    same = (got_ids.cpu() == ref_ids).float().mean().item()
    assert same == 1.0 or same > 0.98  # near-ties inside fp32 rounding may swap neighbours
    # invalid (already seen) ids never come back
    seen = row["historical_ids"]
    for b in range(9):
        assert not set(got_ids[b].cpu().tolist()) & (set(seen[b].tolist()) - {0})


def test_cuda_graph_layer_stack_matches_eager():
    """The captured (forward, backward) graphs over zero-padded jagged rows give the same loss and
    gradients as the eager path, for two different batches replaying the same graph."""
    cfg = RetrievalConfig(name="g", num_items=500, max_sequence_length=40, gr_output_length=5,
                          embedding_dim=128, num_blocks=2, num_heads=2, attention_dim=64,
                          linear_dim=64, dropout=0.0, sampler="local", num_negatives=16, top_k=20,
                          split_year_embedding=False, compute_dtype=torch.bfloat16)
    ids = synthetic_item_ids(300, cfg.num_items, seed=1)
    torch.manual_seed(0)
    m = RetrievalModel(cfg, ids).to(DEV).train()
    rows = [synthetic_batch(cfg, ids, 6, seed=s, min_len=2) for s in (3, 4, 5)]
    draws = {}

    def run(graphs: bool):
        if graphs:
            m.enable_cuda_graphs(row_granularity=256)
        else:
            m.disable_cuda_graphs()
        out = []
        for i, row in enumerate(rows):
            n_rows = int(row["history_lengths"].sum())
            if i not in draws:
                draws[i] = torch.randint(0, ids.numel(), (n_rows, cfg.num_negatives), device=DEV)
            m.negatives_sampler._draw = (lambda d: lambda p, n: m.negatives_sampler._all_item_ids[d])(draws[i])
            m.zero_grad(set_to_none=True)
            loss = m.training_loss({k: v.clone() for k, v in row.items()}, total_length=n_rows)
            loss.backward()
            out.append((loss.item(), {k: p.grad.clone() for k, p in m.named_parameters()
                                      if p.grad is not None}))
        return out

    eager = run(False)
    graphed = run(True)
    assert len(m.sequence_encoder._hstu._graphs) == 1          # one bucket serves all three
    for (le, ge), (lg, gg) in zip(eager, graphed):
        assert abs(le - lg) <= 1e-6 * abs(le)
        assert ge.keys() == gg.keys()
        for k in ge:
            scale = max(ge[k].abs().max().item(), 1e-8)
            # identical kernels; only the fp32 atomics (bias / embedding gradients) reorder
            assert (ge[k] - gg[k]).abs().max().item() <= 2e-3 * scale, k


@pytest.mark.parametrize("sampler", ["inbatch", "local"])
def test_whole_step_cuda_graph_matches_eager(sampler):
    """enable_step_graphs: embedding lookup, sampler, encoder, loss and their backward as one
    captured graph pair over zero-padded rows == the eager, unpadded step (same draws)."""
    cfg = RetrievalConfig(name="g", num_items=500, max_sequence_length=40, gr_output_length=5,
                          embedding_dim=128, num_blocks=2, num_heads=2, attention_dim=64,
                          linear_dim=64, dropout=0.0, sampler=sampler, num_negatives=16, top_k=20,
                          split_year_embedding=False, compute_dtype=torch.bfloat16)
    ids = synthetic_item_ids(300, cfg.num_items, seed=1)
    torch.manual_seed(0)
    m = RetrievalModel(cfg, ids).to(DEV).train()
    with torch.no_grad():
        for p in m.embeddings.parameters():
            p.mul_(10.0)
    rows = [synthetic_batch(cfg, ids, 6, seed=s, min_len=2) for s in (3, 4, 5)]
    fixed_raw = torch.randint(0, 2 ** 40, (1024, cfg.num_negatives), device=DEV)
    smp = m.negatives_sampler

    def draw(positive_ids, n):
        if sampler == "local":
            return smp._all_item_ids[fixed_raw[: positive_ids.size(0)] % smp._all_item_ids.numel()]
        count = smp._cached_count if smp._cached_count is not None else smp._cached_ids.size(0)
        return fixed_raw[: positive_ids.size(0)] % count
    smp._draw = draw

    def run(graphs: bool):
        if graphs:
            m.enable_step_graphs(row_granularity=256)
        else:
            m.disable_cuda_graphs()
        out = []
        for row in rows:
            n_rows = int(row["history_lengths"].sum())
            m.zero_grad(set_to_none=True)
            loss = m.training_loss({k: v.clone() for k, v in row.items()}, total_length=n_rows)
            loss.backward()
            out.append((loss.item(), {k: p.grad.clone() for k, p in m.named_parameters()
                                      if p.grad is not None}))
        return out

    eager = run(False)
    graphed = run(True)
    assert len(m._step_graphs) == 1                  # one bucket serves all three batches
    for (le, ge), (lg, gg) in zip(eager, graphed):
        assert abs(le - lg) <= 1e-5 * abs(le)
        assert ge.keys() == gg.keys()
        for k in ge:
            scale = max(ge[k].abs().max().item(), 1e-8)
            assert (ge[k] - gg[k]).abs().max().item() <= 2e-3 * scale, k


def test_reference_module_code_runs_on_the_gpu_through_the_installed_operator_layer():
    """b-1 under the reference's OWN module code: the unmodified ``HSTU`` of the staged reference package
    (oracle/_ref) on the GPU, its ``ops.<fn>`` call sites (hstu.py:179-204, 502, 512, 661) rebound by
    ``install_ops()`` to this package's kernels, against the same module on the CPU with the stock Python
    fallbacks: forward and every gradient (fp32, 1e-4 relative to the largest entry)."""
    import copy
    import logging
    import sys
    from oracle import ref_verbatim as RV
    if not RV.available():
        pytest.skip("oracle/_ref is not staged")
    sys.path.insert(0, str(RV.REF_ROOT))
    logging.disable(logging.CRITICAL)
    from generative_recommenders_pl.models.sequential_encoders.hstu import HSTU as RefHSTU
    from mygenerativerecommenders_b200 import install
    try:
        torch.manual_seed(5)
        B, L, out_len, D = 6, 40, 5, 32
        N = L + out_len
        enc = RefHSTU(max_sequence_len=L, max_output_len=out_len, embedding_dim=D, item_embedding_dim=D,
                      num_blocks=2, num_heads=2, linear_dim=16, attention_dim=16, normalization="rel_bias",
                      linear_config="uvqk", linear_activation="silu", linear_dropout_rate=0.0,
                      attn_dropout_rate=0.0).eval()
        lengths = torch.tensor([45, 3, 17, 30, 1, 22])
        ts = 978_300_000 + torch.cumsum(torch.randint(1, 5000, (B, N)), dim=1)
        valid = torch.arange(N).unsqueeze(0) < lengths.unsqueeze(1)
        ts = ts * valid
        x = torch.randn(B, N, D) * valid.unsqueeze(-1)
        xc = x.clone().requires_grad_(True)
        yc, _ = enc(past_lengths=lengths, user_embeddings=xc, valid_mask=valid.unsqueeze(-1).float(),
                    past_payloads={"timestamps": ts})
        w = torch.randn_like(yc)
        (yc * w).sum().backward()
        cpu_grads = {k: p.grad.clone() for k, p in enc.named_parameters()}
        install.install_ops()
        encg = copy.deepcopy(enc).to(DEV)
        encg.zero_grad(set_to_none=True)
        xg = x.to(DEV).requires_grad_(True)
        yg, _ = encg(past_lengths=lengths.to(DEV), user_embeddings=xg,
                     valid_mask=valid.unsqueeze(-1).float().to(DEV), past_payloads={"timestamps": ts.to(DEV)})
        (yg * w.to(DEV)).sum().backward()
        assert (yg.cpu() - yc).abs().max().item() <= 1e-4 * yc.abs().max().item()
        assert (xg.grad.cpu() - xc.grad).abs().max().item() <= 1e-4 * xc.grad.abs().max().item()
        for k, p in encg.named_parameters():
            assert (p.grad.cpu() - cpu_grads[k]).abs().max().item() <= 1e-4 * cpu_grads[k].abs().max().item(), k
    finally:
        install.uninstall()
        logging.disable(logging.NOTSET)
        sys.path.remove(str(RV.REF_ROOT))
