"""Drop-in checks that need no GPU: the Hydra override targets resolve, constructors accept the
reference's kwargs, install()/uninstall() rebind the reference's operator layer (when the
reference tree is present — the build container only), and the world_size-2 merge logic."""
import importlib
import os
import sys
from pathlib import Path

import pytest
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference/src")


def _targets(node, out):
    if isinstance(node, dict):
        for k, v in node.items():
            if k == "_target_":
                out.append(v)
            else:
                _targets(v, out)
    return out


def test_override_yaml_targets_resolve():
    for name in ["hstu_b200.yaml", "hstu_b200_bf16.yaml"]:
        cfg = yaml.safe_load((ROOT / "configs" / "model" / name).read_text())
        for t in _targets(cfg, []):
            if t.startswith("hydra."):
                continue
            mod, cls = t.rsplit(".", 1)
            assert hasattr(importlib.import_module(mod), cls), t
    cfg = yaml.safe_load((ROOT / "configs" / "experiment" / "ml-20m-hstu-b200.yaml").read_text())
    assert cfg["model"]["sequence_encoder"]["num_heads"] == 4


def test_constructors_accept_reference_kwargs():
    from mygenerativerecommenders_b200 import candidate_index, hstu, losses, negative_sampler, top_k
    # kwargs exactly as generative_recommenders.py:144-211 / configs/model/hstu.yaml pass them
    enc = hstu.HSTU(max_sequence_len=200, max_output_len=11, embedding_dim=50, item_embedding_dim=50,
                    num_blocks=2, num_heads=1, attention_dim=50, linear_dim=50,
                    linear_dropout_rate=0.2, attn_dropout_rate=0.0, normalization="rel_bias",
                    linear_config="uvqk", linear_activation="silu", concat_ua=False,
                    enable_relative_attention_bias=True)
    assert enc.debug_str() == "HSTU-b2-h1-dqk50-dv50-lsilud0.2-ad0.0"
    smp = negative_sampler.LocalNegativesSampler(l2_norm=True, l2_norm_eps=1e-6, all_item_ids=[1, 2, 5])
    assert smp.debug_str() == "local-l2-eps1e-06" and smp._all_item_ids.tolist() == [1, 2, 5]
    with pytest.raises(ValueError):
        negative_sampler.LocalNegativesSampler(l2_norm=True, l2_norm_eps=1e-6)
    ib = negative_sampler.InBatchNegativesSampler(l2_norm=True, l2_norm_eps=1e-6, dedup_embeddings=True)
    assert ib.debug_str() == "in-batch-l2-eps1e-06-dedup"
    ci = candidate_index.CandidateIndex(k=200, ids=torch.arange(1, 51), top_k_module=top_k.MIPSBruteForceTopK())
    assert ci.num_objects == 50 and ci._k == 50 and ci.ids.shape == (1, 50) and ci.embeddings is None
    ci.update_embeddings(torch.randn(1, 50, 8))
    assert ci._embeddings_t.shape == (8, 50) and ci.embeddings.shape == (1, 50, 8)
    assert losses.SampledSoftmaxLoss(num_to_sample=128, softmax_temperature=0.05)._num_to_sample == 128


@pytest.mark.skipif(not REF_SRC.exists(), reason="reference tree only exists in the build container")
def test_install_rebinds_the_reference_operator_layer():
    sys.path.insert(0, str(REF_SRC))
    import logging
    logging.disable(logging.CRITICAL)
    try:
        from generative_recommenders_pl.models.utils import ops as ref_ops
        from generative_recommenders_pl.models.sequential_encoders import hstu as ref_hstu
        from mygenerativerecommenders_b200 import install, ops as our_ops, hstu as our_hstu
        original = ref_ops.dense_to_jagged
        install.install()
        assert ref_ops.dense_to_jagged is our_ops.dense_to_jagged
        assert ref_ops.asynchronous_complete_cumsum is our_ops.asynchronous_complete_cumsum
        assert ref_hstu.HSTU is our_hstu.HSTU
        # the reference's own call sites now reach our kernels: a CPU tensor must be refused
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ref_ops.mask_dense_by_aux_mask(torch.zeros(1, 2, 1), torch.ones(1, 2, dtype=torch.bool),
                                           torch.tensor([2]), 2)
        install.uninstall()
        assert ref_ops.dense_to_jagged is original
    finally:
        logging.disable(logging.NOTSET)
        sys.path.remove(str(REF_SRC))


def _merge_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mygenerativerecommenders_b200 import candidate_index as ci
    from oracle import reference_port as O
    # the selection kernel is CUDA-only: on CPU the gather / pad logic is exercised with the
    # oracle's selection (same contract: sorted, ties -> lowest id)
    def cpu_select(s, i, k):
        order = torch.argsort(i, dim=1, stable=True)
        s, i = torch.gather(s, 1, order), torch.gather(i, 1, order)
        order = torch.argsort(s, dim=1, descending=True, stable=True)[:, :k]
        return torch.gather(s, 1, order), torch.gather(i, 1, order)
    ci.GF.topk_merge = cpu_select
    g = torch.Generator().manual_seed(0)
    X, D, B, k = 1000, 16, 7, 40
    items = torch.randn(X, D, generator=g)
    q = torch.randn(B, D, generator=g)
    per = -(-X // world) if rank == 0 else -(-X // world)
    lo, hi = rank * per, min((rank + 1) * per, X)
    # rank 1 deliberately holds fewer than k items to exercise the padding path
    if rank == 1:
        hi = lo + 25
    s, i = O.mips_topk(q, items[lo:hi], torch.arange(lo, hi), min(k, hi - lo))
    ms, mi = ci.merge_sharded_topk(s, i, k, world)
    union = torch.cat([torch.arange(0, per), torch.arange(per, per + 25)])
    rs, ri = O.mips_topk(q, items[union], union, k)
    assert torch.equal(mi, ri) and torch.allclose(ms, rs)
    torch.save(mi, f"{tmp}/merged_{rank}.pt")
    dist.destroy_process_group()


def test_sharded_topk_merge_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_merge_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "merged_0.pt"), torch.load(tmp_path / "merged_1.pt")
    assert torch.equal(a, b)  # every rank ends with the same merged top-k


def test_samplers_pass_the_reference_trainers_isinstance_check():
    """models/retrieval.py:104 picks process_batch vs. the late-bound embeddings module with
    isinstance(..., InBatchNegativesSampler) against the REFERENCE class; the drop-in samplers
    derive from the reference classes whenever its package is importable."""
    import importlib
    import sys
    ref_src = "/root/reference/src"
    import os
    if not os.path.isdir(ref_src):
        pytest.skip("no reference checkout in this environment")
    sys.path.insert(0, ref_src)
    try:
        ref = importlib.import_module("generative_recommenders_pl.models.negatives_samples.negative_sampler")
        import mygenerativerecommenders_b200.negative_sampler as ours
        ours = importlib.reload(ours)        # pick up the reference bases now that it is importable
        ib = ours.InBatchNegativesSampler(l2_norm=True, l2_norm_eps=1e-6, dedup_embeddings=True)
        lo = ours.LocalNegativesSampler(l2_norm=True, l2_norm_eps=1e-6, all_item_ids=[1, 2, 3])
        assert isinstance(ib, ref.InBatchNegativesSampler) and isinstance(ib, ref.NegativesSampler)
        assert isinstance(lo, ref.LocalNegativesSampler) and not isinstance(lo, ref.InBatchNegativesSampler)
        assert ib._dedup_embeddings and lo._num_items == 3 and hasattr(ib, "fused_sample")
    finally:
        sys.path.remove(ref_src)
        import mygenerativerecommenders_b200.negative_sampler as ours2
        importlib.reload(ours2)
