"""The CPU oracle (oracle/reference_port.py) against fixtures produced by the real reference
(oracle/make_golden.py).  CPU-only; pins the oracle before any kernel is compared with it."""
import pytest
import torch

from oracle import reference_port as O
from conftest import hstu_case, hstu_incremental_case


def test_ops_known_answers(golden):
    g = golden("ops")
    # the reference's own vectors, /root/reference/tests/test_ops.py:7-53
    assert torch.equal(O.complete_cumsum(g["kat_cumsum_in"]), g["kat_cumsum_out"])
    assert O.complete_cumsum(g["kat_cumsum_in"]).dtype == torch.int32
    assert torch.equal(g["kat_cumsum_out"], torch.tensor([0, 1, 3], dtype=torch.int32))
    assert torch.equal(O.dense_to_jagged(g["kat_d2j_in"], g["kat_cumsum_out"]), g["kat_d2j_out"])
    assert torch.equal(O.jagged_to_padded_dense(g["kat_j2d_in"], g["kat_j2d_off"], 3, 0.0),
                       g["kat_j2d_out"])


def test_ops_ragged(golden):
    g = golden("ops")
    off = O.complete_cumsum(g["r_lengths"])
    assert torch.equal(off, g["r_offsets"])
    jag = O.dense_to_jagged(g["r_dense"], off)
    assert torch.equal(jag, g["r_jagged"])
    assert torch.equal(O.jagged_to_padded_dense(jag, off, 12, 0.0), g["r_padded"])
    assert torch.equal(O.jagged_to_padded_dense(jag, off, 12, -1.5), g["r_padded_pad"])
    assert torch.equal(O.get_current_embeddings(g["cur_lengths"], g["r_dense"]), g["cur_out"])
    out, nl = O.mask_dense_by_aux_mask(g["mask_dense"], g["mask_aux"], g["mask_lengths"], 4)
    assert torch.equal(out, g["mask_out"]) and torch.equal(nl, g["mask_new_lengths"])


def test_bucket_and_bias(golden):
    g = golden("bias")
    assert torch.equal(O.bucketize_ts(g["bucket_probe"]), g["bucket_value"])
    N = g["bias_ts"].shape[1]
    out = O.rel_bias(g["bias_ts"], g["bias_ts_w"], g["bias_pos_w"], N)
    assert torch.equal(out, g["bias_out"])  # same ops in the same order: bit-exact


@pytest.mark.parametrize("name", ["mh", "ml1m", "h64"])
def test_hstu_forward_backward(golden, name):
    c = hstu_case(golden("hstu"), name)
    sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
    x = c["x"].clone().requires_grad_(True)
    y = O.hstu_forward(c["lengths"], x, c["ts"], sd, c["blocks"], c["H"], c["dqk"], c["dv"])
    # fp32 on both sides, same algorithm: tolerance 1e-5 relative to the output scale
    scale = c["y"].abs().max().item()
    assert (y - c["y"]).abs().max().item() <= 1e-5 * scale
    (y * c["w"]).sum().backward()
    assert (x.grad - c["dx"]).abs().max().item() <= 1e-4 * c["dx"].abs().max().item()
    for k, gref in c["grads"].items():
        got = sd[k].grad
        assert got is not None, k
        assert (got - gref).abs().max().item() <= 2e-4 * max(gref.abs().max().item(), 1e-6), k


def test_topk_and_candidate_index(golden):
    g = golden("retrieval")
    s, i = O.mips_topk(g["tk_q"], g["tk_table"], g["tk_ids"], 25)
    assert torch.equal(i, g["tk_ids25"])
    assert torch.allclose(s, g["tk_scores25"], atol=1e-6)
    oi, os_ = O.candidate_index_topk(g["tk_q"], g["tk_table"], g["tk_ids"], 10, g["ci_invalid"])
    assert torch.equal(oi, g["ci_ids"])
    assert torch.allclose(os_, g["ci_scores"], atol=1e-6)


def test_topk_tie_rule():
    # duplicated items => equal scores; the lowest index must win, in order
    q = torch.tensor([[1.0, 0.0]])
    items = torch.tensor([[0.5, 0.0], [0.9, 0.0], [0.9, 0.0], [0.1, 0.0], [0.9, 0.0]])
    s, i = O.mips_topk(q, items, None, 3)
    assert i.tolist() == [[1, 2, 4]]
    s, i = O.mips_topk(q, items, None, 2, chunk=2)
    assert i.tolist() == [[1, 2]]


def test_sampled_softmax(golden):
    g = golden("retrieval")
    table = g["ssl_table"].clone().requires_grad_(True)
    out_emb = g["ssl_out_emb"].clone().requires_grad_(True)
    neg_emb = table[g["ssl_neg_ids"]]
    sup_emb = table[g["ssl_sup_ids"]]
    loss, _ = O.sampled_softmax_loss(out_emb, g["ssl_sup_ids"], sup_emb, g["ssl_sup_w"],
                                     g["ssl_neg_ids"], neg_emb, 0.05, 1e-6)
    assert abs(loss.item() - g["ssl_loss"].item()) <= 1e-5 * abs(g["ssl_loss"].item())
    loss.backward()
    assert torch.allclose(out_emb.grad, g["ssl_d_out_emb"], rtol=1e-4, atol=1e-6)
    dt = table.grad.clone()
    dt[0] = 0  # padding_idx row of the reference's nn.Embedding
    assert torch.allclose(dt, g["ssl_d_table"], rtol=1e-4, atol=1e-6)


def test_inbatch_dedup(golden):
    g = golden("retrieval")
    ids, emb = O.inbatch_process(g["ib_ids"], g["ib_ids"] != 0, g["ib_emb"], 1e-6, True)
    # unique(sorted=False) order is unspecified: compare as id -> embedding maps
    ref = {int(i): e for i, e in zip(g["ib_cached_ids"], g["ib_cached_emb"])}
    assert sorted(ref) == sorted(int(i) for i in ids)
    for i, e in zip(ids, emb):
        assert torch.allclose(e, ref[int(i)], atol=1e-7)


@pytest.mark.parametrize("name", ["mh", "h64"])
def test_hstu_incremental_path(golden, name):
    """delta_x_offsets + cache (hstu.py:293-298, :151-177, :415-418): cache states of a full pass,
    then the last token of every sequence recomputed from them, against the real reference."""
    c = hstu_incremental_case(golden("hstu_incremental"), name)
    args = (c["sd"], c["blocks"], c["H"], c["dqk"], c["dv"])
    states = O.hstu_cache_states(c["lengths"], c["x"], c["ts"], *args)
    for got, ref in zip(states, c["cache0"]):
        for a, b in zip(got, ref):
            assert a.shape == b.shape
            assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-6)
    y, new_states = O.hstu_forward_incremental(c["lengths"], c["x2"], c["ts"], *args, c["delta"], c["cache0"])
    assert (y - c["y_inc"]).abs().max().item() <= 1e-5 * c["y_inc"].abs().max().item()
    # only the recomputed rows differ from the first pass
    changed = ((y - c["y0"]).abs().sum(-1) > 0)
    last = torch.zeros_like(changed)
    last[torch.arange(c["B"]), c["lengths"] - 1] = True
    assert torch.equal(changed & ~last, torch.zeros_like(changed))
    if c["cache1"] is not None:
        for got, ref in zip(new_states, c["cache1"]):
            for a, b in zip(got, ref):
                assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-6)


@pytest.mark.parametrize("name", ["mh", "h64"])
def test_hstu_softmax_rel_bias(golden, name):
    """normalization="softmax_rel_bias" (hstu.py:337-384) against the real reference."""
    c = hstu_case(golden("hstu_softmax"), name)
    sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
    x = c["x"].clone().requires_grad_(True)
    y = O.hstu_forward(c["lengths"], x, c["ts"], sd, c["blocks"], c["H"], c["dqk"], c["dv"],
                       normalization="softmax_rel_bias")
    assert (y - c["y"]).abs().max().item() <= 1e-5 * c["y"].abs().max().item()
    (y * c["w"]).sum().backward()
    assert (x.grad - c["dx"]).abs().max().item() <= 1e-4 * c["dx"].abs().max().item()
    for k, gref in c["grads"].items():
        assert (sd[k].grad - gref).abs().max().item() <= 2e-4 * max(gref.abs().max().item(), 1e-6), k


OPTION_CASES = {   # name -> (oracle kwargs, timestamps passed?)
    "ua": (dict(concat_ua=True), True), "ua64": (dict(concat_ua=True), True),
    "noact": (dict(linear_activation="none"), True), "norab": ({}, False),
}


@pytest.mark.parametrize("name", sorted(OPTION_CASES))
def test_hstu_layer_options(golden, name):
    """concat_ua (hstu.py:398-402), linear_activation="none" (:304-307), no relative bias."""
    c = hstu_case(golden("hstu_options"), name)
    kw, with_ts = OPTION_CASES[name]
    sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
    x = c["x"].clone().requires_grad_(True)
    y = O.hstu_forward(c["lengths"], x, c["ts"] if with_ts else None, sd, c["blocks"], c["H"], c["dqk"],
                       c["dv"], **kw)
    assert (y - c["y"]).abs().max().item() <= 1e-5 * c["y"].abs().max().item()
    (y * c["w"]).sum().backward()
    assert (x.grad - c["dx"]).abs().max().item() <= 1e-4 * c["dx"].abs().max().item()
    for k, gref in c["grads"].items():
        assert (sd[k].grad - gref).abs().max().item() <= 2e-4 * max(gref.abs().max().item(), 1e-6), k


def test_port_agrees_with_the_staged_reference_modules_live():
    """Beyond the committed fixtures: where the unmodified reference package is staged (oracle/_ref, written
    by oracle/stage_ref.py in the build container), the port's encoder and its sampled-softmax loss are run
    against the reference's OWN modules on fresh random weights and inputs (CPU, fp32)."""
    from oracle import ref_verbatim as RV
    if not RV.available():
        pytest.skip("oracle/_ref is not staged")
    R = RV._import()
    torch.manual_seed(11)
    B, L, out_len, D, H, d = 5, 30, 4, 48, 2, 24
    N = L + out_len
    enc = R["HSTU"](max_sequence_len=L, max_output_len=out_len, embedding_dim=D, item_embedding_dim=D,
                    num_blocks=2, num_heads=H, linear_dim=d, attention_dim=d, normalization="rel_bias",
                    linear_config="uvqk", linear_activation="silu", linear_dropout_rate=0.0,
                    attn_dropout_rate=0.0).eval()
    lengths = torch.tensor([34, 2, 17, 30, 9])
    valid = torch.arange(N).unsqueeze(0) < lengths.unsqueeze(1)
    ts = (978_300_000 + torch.cumsum(torch.randint(1, 5000, (B, N)), dim=1)) * valid
    x = torch.randn(B, N, D) * valid.unsqueeze(-1)
    with torch.no_grad():
        y_ref, _ = enc(past_lengths=lengths, user_embeddings=x, valid_mask=valid.unsqueeze(-1).float(),
                       past_payloads={"timestamps": ts})
        sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
        y = O.hstu_forward(lengths, x, ts, sd, 2, H, d, d)
    assert (y - y_ref).abs().max().item() <= 1e-5 * y_ref.abs().max().item()
    # sampled softmax: the reference's loss with its local sampler's draw injected into the port
    V, R_ = 60, 16
    table = torch.randn(V + 1, D)
    table[0] = 0
    emb = torch.nn.Embedding(V + 1, D, padding_idx=0)
    with torch.no_grad():
        emb.weight.copy_(table)
    smp = R["LocalNegativesSampler"](l2_norm=True, l2_norm_eps=1e-6, all_item_ids=list(range(1, V + 1)))
    smp._item_emb = emb
    n = 40
    out_emb = torch.nn.functional.normalize(torch.randn(n, D), dim=-1)
    sup_ids = torch.randint(1, V + 1, (n,))
    sup_w = (torch.rand(n) > 0.2).float()
    loss_fn = R["SampledSoftmaxLoss"](num_to_sample=R_, softmax_temperature=0.05)
    torch.manual_seed(3)
    with torch.no_grad():
        ref_loss = loss_fn.jagged_forward(output_embeddings=out_emb, supervision_ids=sup_ids,
                                          supervision_embeddings=table[sup_ids], supervision_weights=sup_w,
                                          negatives_sampler=smp, similarity=R["DotProductSimilarity"]())
    torch.manual_seed(3)
    neg_ids = smp._all_item_ids[torch.randint(low=0, high=V, size=(n, R_))] if hasattr(smp, "_all_item_ids") else None
    if neg_ids is None:
        pytest.skip("reference sampler layout changed")
    loss, _ = O.sampled_softmax_loss(out_emb, sup_ids, table[sup_ids], sup_w, neg_ids, table[neg_ids], 0.05, 1e-6)
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
