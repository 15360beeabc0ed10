"""CPU-side checks: the C-ABI library loads and exports what include/grb200.h declares, the host
modules mirror the reference's names, and the product path refuses CPU tensors (no fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

import mygenerativerecommenders_b200 as pkg
from mygenerativerecommenders_b200 import _lib, hstu, ops
from conftest import hstu_case

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "grb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 15
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(handle, n), f"libgrb200.so lacks {n}"
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    assert _lib.lib().grb_version() == 100


def test_struct_layouts_match_header():
    # sizes implied by the header's field lists (8-byte aligned, LP64)
    assert ctypes.sizeof(_lib.HstuAttnArgs) == 4 * 8 + 6 * 4 + 3 * 8 + 3 * 8 + 5 * 8 + 2 * 8 + 2 * 8 + 3 * 8 + 3 * 8 + 3 * 8 + 8 + 8 + 16 + 16
    assert ctypes.sizeof(_lib.MipsTopkArgs) == 3 * 8 + 2 * 4 + 4 * 8 + 8 + 2 * 8 + 2 * 8 + 2 * 8 + 8 + (8 + 8 + 8) + 8 + 8
    assert ctypes.sizeof(_lib.SslArgs) == 8 + 6 * 4 + 2 * 4 + 8 * 8 + 4 * 8 + 2 * 8 + 5 * 8
    # grb_ln_gate_args: 3 x (ptr, ld) + 2 ptr + rows, W + (eps, dtype) + (p_drop, pad) + seed, salt + 3 x (ptr, ld)
    assert ctypes.sizeof(_lib.LnGateArgs) == 6 * 8 + 2 * 8 + 2 * 8 + 8 + 8 + 2 * 8 + 6 * 8
    assert _lib.LnGateArgs.seed.offset == 96 and _lib.LnGateArgs.dx.offset == 112


def test_invalid_arguments_are_reported_without_a_gpu():
    L = _lib.lib()
    assert L.grb_complete_cumsum(None, None, 4, 16, None) == _lib.GRB_ERR_INVALID_ARG
    assert b"index_bits" in L.grb_last_error_string()
    with pytest.raises(ValueError):
        _lib.check(L.grb_complete_cumsum(None, None, 4, 16, None))
    a = _lib.MipsTopkArgs()
    a.B, a.X, a.D, a.k = 4, 100, 8, 200  # k > X
    assert L.grb_mips_topk_workspace_bytes(ctypes.byref(a)) == _lib.GRB_ERR_INVALID_ARG


def _mips_ws(B, X, D, k, dtype, n_invalid=0, cand_cap=0):
    a = _lib.MipsTopkArgs()
    a.B, a.X, a.D, a.k, a.dtype = B, X, D, k, dtype
    a.ldq = a.ldi = D
    a.n_invalid, a.cand_cap = n_invalid, cand_cap
    if n_invalid:
        a.invalid_ids = 16          # any non-null address: the planner only looks at the count
    need = int(_lib.lib().grb_mips_topk_workspace_bytes(ctypes.byref(a)))
    return need, int(a.cand_cap), int(a.sample_stride)


def test_topk_planner_picks_the_one_query_block_plan_where_it_applies():
    """Host logic of grb_mips_topk_workspace_bytes (no GPU): one bf16 query block against a corpus of several
    hundred tiles gets the workspace of the one-query-block plan (csrc/mips_small.cu: private sub-lists, much
    larger than the phased plan's candidate lists); more queries, fp32 tables, tiny corpora, or an explicit
    candidate capacity (the wrapper's exact re-run after an overflow) get the phased plan's."""
    BF16, F32 = _lib.GRB_BF16, _lib.GRB_F32
    c3, cap, stride = _mips_ws(128, 700_000, 64, 200, BF16, n_invalid=61)
    assert stride == 64 and cap == 261 * (4 + 2 * 3 * 3) + 1024       # the phased plan's own sizes are still reported
    rerun, cap2, _ = _mips_ws(128, 700_000, 64, 200, BF16, n_invalid=61, cand_cap=cap + 1024)
    assert cap2 == cap + 1024
    phased_c3 = 128 * (4 + 4 + 11008 * 4 + 2 * 4 * cap)               # tau, counts, sample, candidate lists
    assert rerun < 2 * phased_c3 + 128 * 2 * 4 * 1024 and c3 > 2 * rerun   # sub-lists: 128 x 592 x (24 + 32) x 8 B
    assert abs(c3 - 128 * 592 * 56 * 8) < 8 << 20
    two_blocks, _, _ = _mips_ws(129, 700_000, 64, 200, BF16, n_invalid=61)
    fp32, _, _ = _mips_ws(128, 700_000, 64, 200, F32, n_invalid=61)
    assert two_blocks < c3 / 2 and fp32 < c3 / 2
    small_corpus, cap_s, _ = _mips_ws(128, 20_000, 64, 200, BF16)      # 157 tiles: too few groups for the estimate
    assert cap_s == 200 * (4 + 2 * 3 * 1) + 1024 and small_corpus < 128 * (8 + 40 * 128 * 4 + 8 * cap_s) + 4096
    # k' too large for the 8192-key select of the plan: phased
    big_k, _, _ = _mips_ws(128, 700_000, 64, 1500, BF16)
    assert big_k < 128 * 2 * 4 * (1500 * 22 + 1024) + (16 << 20)
    # C4 corpus at one query block: sample stride 16, still the one-query-block plan
    c4, _, _ = _mips_ws(128, 10_000_000, 256, 200, BF16)
    assert c4 > 128 * 592 * 40 * 8


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.asynchronous_complete_cumsum(torch.tensor([1, 2]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.dense_to_jagged(torch.zeros(2, 3, 1), torch.tensor([0, 1, 3]))
    with pytest.raises(ValueError, match="max_lengths must be an integer"):
        ops.jagged_to_padded_dense(torch.zeros(3, 1), torch.tensor([0, 1, 3]), [3], 0.0)
    # every functional entry point, including the ones whose body is a library composite
    from mygenerativerecommenders_b200 import functional as GF
    from mygenerativerecommenders_b200.optim import FusedAdamW
    z = torch.zeros(4, 8)
    off, pos = torch.tensor([0, 4]), torch.tensor([3])
    for call in (
        lambda: GF.silu_split(z, [4, 4]),
        lambda: GF.master_linear(z, torch.zeros(8, 8), None, w_in_out=True),
        lambda: GF.linear_bias(z, torch.zeros(8, 8), torch.zeros(8)),
        lambda: GF.l2_normalize(z, 1e-6),
        lambda: GF.layer_norm_gate(z, None, 1e-6),
        lambda: GF.embedding_lookup(torch.zeros(5, 8), torch.tensor([1, 2]), 0),
        lambda: GF.hstu_attention(z, z, z, off, None, None, None, None, 4, 1, 8, 8),
        lambda: GF.hstu_attention_decode(z[:1], z.view(1, 4, 8), z, off, pos, None, None, None, None, 4, 1, 8, 8),
        lambda: GF.mips_topk(z, z, None, 2),
    ):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        FusedAdamW([p]).step()


def test_product_never_imports_oracle():
    for p in (ROOT / "mygenerativerecommenders_b200").rglob("*.py"):
        assert "oracle" not in p.read_text(), f"{p} mentions the oracle"


def test_bucket_threshold_table_matches_reference(golden):
    g = golden("bias")
    thr = hstu.tabulate_bucket_thresholds(hstu._default_bucketization, 128)
    assert thr.dtype == torch.int64 and thr.numel() == 128
    assert (thr[1:] >= thr[:-1]).all()
    got = torch.bucketize(g["bucket_probe"], thr, right=True)
    assert torch.equal(got, g["bucket_value"])
    # first edges are ceil(e^{0.301 k})
    assert thr[:8].tolist() == [2, 2, 3, 4, 5, 7, 9, 12]


def test_dense_bias_module_matches_reference(golden):
    g = golden("bias")
    N = g["bias_ts"].shape[1]
    m = hstu.RelativeBucketedTimeAndPositionBasedBias(N, 128, hstu._default_bucketization)
    with torch.no_grad():
        m._ts_w.copy_(g["bias_ts_w"]); m._pos_w.copy_(g["bias_pos_w"])
    assert torch.equal(m(g["bias_ts"]).detach(), g["bias_out"])


@pytest.mark.parametrize("name", ["mh", "ml1m"])
def test_module_parameter_names_match_reference(golden, name):
    c = hstu_case(golden("hstu"), name)
    enc = hstu.HSTU(max_sequence_len=c["max_seq"], max_output_len=c["out_len"],
                    embedding_dim=c["D"], item_embedding_dim=c["D"], num_blocks=c["blocks"],
                    num_heads=c["H"], linear_dim=c["dv"], attention_dim=c["dqk"],
                    normalization="rel_bias", linear_config="uvqk", linear_activation="silu",
                    linear_dropout_rate=0.2, attn_dropout_rate=0.0)
    ours = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in c["sd"].items()}
    ref["_attn_mask"] = (c["N"], c["N"])
    assert ours == ref
    missing, unexpected = enc.load_state_dict(c["sd"], strict=False)
    assert unexpected == [] and missing == ["_attn_mask"]


def test_public_api_names():
    for n in ["asynchronous_complete_cumsum", "dense_to_jagged", "jagged_to_padded_dense",
              "batch_gather_embeddings", "batch_scatter_embeddings", "get_current_embeddings",
              "jagged_or_dense_repeat_interleave_dim0", "jagged_or_dense_index_select_dim0",
              "mask_dense_by_aux_mask"]:
        assert callable(getattr(ops, n))
    from mygenerativerecommenders_b200 import (candidate_index, losses, negative_sampler,
                                                similarity, top_k)
    assert issubclass(top_k.MIPSBruteForceTopK, top_k.TopKModule)
    assert issubclass(negative_sampler.LocalNegativesSampler, negative_sampler.NegativesSampler)
    assert issubclass(losses.SampledSoftmaxLoss, losses.AutoregressiveLoss)
    assert similarity.DotProductSimilarity().debug_str() == "dp"
    assert candidate_index.CandidateIndex is not None


def test_layer_argument_errors_come_before_any_kernel():
    """Host-side checks of SequentialTransductionUnitJagged.forward (hstu.py:266-423 options): they are
    raised from Python, before any CUDA call, so they can be checked on CPU tensors."""
    mk = lambda **kw: hstu.SequentialTransductionUnitJagged(
        embedding_dim=16, linear_hidden_dim=8, attention_dim=8, dropout_ratio=0.0, attn_dropout_ratio=0.0,
        num_heads=2, linear_activation="silu", **kw)
    x, off, mask = torch.zeros(3, 16), torch.tensor([0, 3]), torch.zeros(4, 4)
    delta = (torch.tensor([2]), torch.tensor([2]))
    with pytest.raises(ValueError, match="cache"):                     # hstu.py:295 `assert cache is not None`
        mk()(x, off, None, mask, delta_x_offsets=delta)
    with pytest.raises(ValueError, match="Unknown normalization"):     # hstu.py:385
        mk(normalization="nope")(x, off, None, mask)
    cache = (torch.zeros(3, 16), torch.zeros(1, 4, 16), torch.zeros(1, 4, 16), torch.zeros(3, 16))
    with pytest.raises(NotImplementedError, match="incremental"):      # the reference raises TypeError, :339
        mk(normalization="softmax_rel_bias")(x, off, None, mask, delta_x_offsets=delta, cache=cache)
    with pytest.raises(RuntimeError, match="no CPU fallback"):         # a valid call reaches the kernels
        mk()(x, off, None, mask)


def test_retrieval_metrics_follow_the_reference_formulas():
    """metrics/retrieval.py:40-68 restated (torchmetrics is not in the image): ranks from (top_k_ids,
    target_ids), then NDCG / HR / MRR; ``update_ranks`` (the fused kernel's output) gives the same."""
    from mygenerativerecommenders_b200.metrics import RetrievalMetrics
    g = torch.Generator().manual_seed(5)
    B, k = 64, 20
    top = torch.stack([torch.randperm(500, generator=g)[:k] + 1 for _ in range(B)])
    tgt = torch.where(torch.rand(B, generator=g) < 0.6, top[torch.arange(B), torch.randint(0, k, (B,), generator=g)],
                      torch.full((B,), 9999))
    m = RetrievalMetrics(k=k, at_k_list=[1, 10, 20])
    m.update(top[:40], tgt[:40].unsqueeze(1))
    m.update(top[40:], tgt[40:])
    out = m.compute()
    # the reference's compute(), line by line
    _, rank_indices = torch.max(torch.cat([top, tgt.unsqueeze(1)], dim=1) == tgt.unsqueeze(1), dim=1)
    ranks = rank_indices + 1
    for at_k in (1, 10, 20):
        ndcg = torch.where(ranks <= at_k, 1.0 / torch.log2(ranks + 1), torch.zeros(1)).mean()
        assert torch.equal(out[f"ndcg@{at_k}"], ndcg)
        assert torch.equal(out[f"hr@{at_k}"], (ranks <= at_k).float().mean())
    assert torch.equal(out["mrr"], (1.0 / ranks).mean())
    assert (ranks == k + 1).any() and (ranks <= k).any()
    m2 = RetrievalMetrics(k=k, at_k_list=[1, 10, 20])
    m2.update_ranks(ranks.to(torch.int32))
    assert all(torch.equal(out[key], v) for key, v in m2.compute().items())


def test_async_topk_handle_reruns_on_overflow_and_gives_up_after_three():
    """functional.MipsTopkCall: result() waits for the call's own flag, re-runs with the reported capacity
    + slack when a row overflowed, returns the tensors of the successful attempt (host logic only)."""
    from mygenerativerecommenders_b200.functional import MipsTopkCall

    class Done:
        def synchronize(self):
            pass

    caps = []

    def run(cap):
        caps.append(cap)
        flag = [0 if cap >= 5000 else 4000]          # the first attempt (cap 0 = auto) overflows to 4000
        return ("scores", "ids", cap), flag, Done()

    call = MipsTopkCall(run, *run(0))
    assert call.result() == ("scores", "ids", 5024) and caps == [0, 5024]
    always = MipsTopkCall(lambda cap: (None, [7], Done()), None, [7], Done())
    with pytest.raises(RuntimeError, match="overflowed repeatedly"):
        always.result()


def test_candidate_index_takes_the_fused_filter_only_when_it_is_equivalent():
    from mygenerativerecommenders_b200.candidate_index import _fused_filter_ok
    from mygenerativerecommenders_b200.top_k import MIPSBruteForceTopK

    class Foreign:      # a reference-style top-k module without forward_filtered
        pass
    ours = MIPSBruteForceTopK()
    assert _fused_filter_ok(ours, 200, 211, 3706)            # C1 / C2: k' = 411
    assert _fused_filter_ok(ours, 200, 61, 700_000)          # C3
    assert not _fused_filter_ok(ours, 200, 0, 3706)          # nothing to filter
    assert not _fused_filter_ok(Foreign(), 200, 61, 3706)    # foreign module: the reference's composite
    assert not _fused_filter_ok(ours, 200, 61, 230)          # k + N exceeds the corpus: rows could run short
    assert not _fused_filter_ok(ours, 1900, 200, 10 ** 6)    # k + N > 2048: beyond the kernel's shared memory
    assert not _fused_filter_ok(ours, 10, 1025, 10 ** 6)     # invalid list longer than the kernel stages
