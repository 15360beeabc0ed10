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
