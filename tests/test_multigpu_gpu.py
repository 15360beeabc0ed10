"""N > 1 on real GPUs (skipped with fewer than 2 devices): corpus-sharded top-k with the
peer-memory exchange + merge kernel must equal the single-GPU result bit for bit on every rank."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mygenerativerecommenders_b200.candidate_index import CandidateIndex, ShardedCandidateIndex
    from mygenerativerecommenders_b200.top_k import MIPSBruteForceTopK
    dev = torch.device("cuda", rank)
    g = torch.Generator().manual_seed(0)
    X, D, B, k = 100_003, 64, 33, 200
    ids = torch.arange(1, X + 1) * 2
    table = torch.nn.functional.normalize(torch.randn(X, D, generator=g), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1).to(torch.bfloat16).to(dev)
    invalid = ids[torch.randint(0, X, (B, 61), generator=g)].to(dev)
    sharded = ShardedCandidateIndex(k=k, ids=ids, top_k_module=MIPSBruteForceTopK(),
                                    embeddings=table.unsqueeze(0).to(dev)).to(dev)
    assert sharded._ids.shape[1] == -(-X // world) or rank == world - 1
    si, ss = sharded.get_top_k_outputs(q, invalid_ids=invalid)
    full = CandidateIndex(k=k, ids=ids, top_k_module=MIPSBruteForceTopK(),
                          embeddings=table.unsqueeze(0).to(dev)).to(dev)
    fi, fs = full.get_top_k_outputs(q, invalid_ids=invalid)
    assert torch.equal(si, fi), (si != fi).sum()
    assert torch.equal(ss, fs)
    # the exchange went over peer memory (symmetric buffers + P2P stores), not NCCL ...
    from mygenerativerecommenders_b200.candidate_index import _PeerExchange
    assert _PeerExchange._cache and not _PeerExchange._broken
    # ... and the buffers are reusable round after round (barrier on both sides of the puts)
    for rep in range(3):
        q2 = torch.roll(q, rep + 1, 0)
        si2, ss2 = sharded.get_top_k_outputs(q2, invalid_ids=invalid)
        fi2, fs2 = full.get_top_k_outputs(q2, invalid_ids=invalid)
        assert torch.equal(si2, fi2) and torch.equal(ss2, fs2)
    torch.save(si.cpu(), f"{tmp}/ids_{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_candidate_index_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, 29400 + os.getpid() % 500, str(tmp_path)), nprocs=world, join=True)
    a, b = torch.load(tmp_path / "ids_0.pt"), torch.load(tmp_path / "ids_1.pt")
    assert torch.equal(a, b)


def _worker_table_grads(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from mygenerativerecommenders_b200.pipeline import (RetrievalConfig, RetrievalModel,
                                                        synthetic_batch, synthetic_item_ids)
    cfg = RetrievalConfig(name="p", num_items=500, max_sequence_length=40, gr_output_length=5,
                          embedding_dim=128, num_blocks=1, num_heads=2, attention_dim=64,
                          linear_dim=64, dropout=0.0, sampler="inbatch", num_negatives=16, top_k=20,
                          split_year_embedding=False, compute_dtype=torch.bfloat16)
    ids = synthetic_item_ids(300, cfg.num_items, seed=1)
    torch.manual_seed(0)
    m = RetrievalModel(cfg, ids).to(dev).train()
    row = synthetic_batch(cfg, ids, 6, seed=10 + rank, min_len=2)      # a different batch per rank
    n_rows = int(row["history_lengths"].sum())
    raw = torch.randint(0, 2 ** 40, (n_rows, cfg.num_negatives), device=dev)
    smp = m.negatives_sampler
    smp._draw = lambda p, n: raw % smp._cached_count
    w = m.embeddings._item_emb.weight
    # reference: dense gradient averaged with NCCL
    m.training_loss({k: v.clone() for k, v in row.items()}, total_length=n_rows).backward()
    ref = w.grad.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    # the same step with the peer-memory exchange, twice (buffers are reused)
    m.enable_peer_table_grads()
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        m.training_loss({k: v.clone() for k, v in row.items()}, total_length=n_rows).backward()
        scale = ref.abs().max().item()
        assert (w.grad - ref).abs().max().item() <= 1e-5 * scale
        assert (w.grad[0] == 0).all()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_memory_table_gradient_equals_allreduce(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker_table_grads, args=(world, 29900 + os.getpid() % 90, str(tmp_path)), nprocs=world,
             join=True)


def _worker_all_grads(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from mygenerativerecommenders_b200.pipeline import (RetrievalConfig, RetrievalModel,
                                                        synthetic_batch, synthetic_item_ids)
    cfg = RetrievalConfig(name="p", num_items=500, max_sequence_length=40, gr_output_length=5,
                          embedding_dim=128, num_blocks=2, num_heads=2, attention_dim=64,
                          linear_dim=64, dropout=0.0, sampler="inbatch", num_negatives=16, top_k=20,
                          split_year_embedding=False, compute_dtype=torch.bfloat16)
    ids = synthetic_item_ids(300, cfg.num_items, seed=1)
    torch.manual_seed(0)
    m = RetrievalModel(cfg, ids).to(dev).train()
    row = synthetic_batch(cfg, ids, 6, seed=10 + rank, min_len=2)      # a different batch per rank
    n_rows = int(row["history_lengths"].sum())
    raw = torch.randint(0, 2 ** 40, (n_rows, cfg.num_negatives), device=dev)
    smp = m.negatives_sampler
    smp._draw = lambda p, n: raw % smp._cached_count
    # reference: every gradient averaged with NCCL
    m.training_loss({k: v.clone() for k, v in row.items()}, total_length=n_rows).backward()
    ref = {}
    for name, p in m.named_parameters():
        g = p.grad.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        ref[name] = g / world
    red = m.enable_peer_gradients()
    for _ in range(3):          # buffers are reused step after step
        m.zero_grad(set_to_none=True)
        m.training_loss({k: v.clone() for k, v in row.items()}, total_length=n_rows).backward()
        red.reduce()
        for name, p in m.named_parameters():
            scale = max(ref[name].abs().max().item(), 1e-30)
            assert (p.grad - ref[name]).abs().max().item() <= 1e-5 * scale, name
    # all ranks hold bit-identical dense gradients (sums run in rank order)
    flat = red.flat.clone()
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    assert all(torch.equal(o, other[0]) for o in other)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_memory_gradient_reduction_equals_allreduce(tmp_path):
    """PeerGradients (sparse table rows + two-shot all-reduce kernel over peer memory, no DDP) against
    NCCL averages of the same per-rank gradients."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker_all_grads, args=(world, 29700 + os.getpid() % 90, str(tmp_path)), nprocs=world, join=True)
