"""a1-a3, a10 on the GPU through the C ABI: bit-exact against the reference's golden vectors,
the CPU oracle on ragged inputs, and size-independent properties at BASELINE sizes."""
import pytest
import torch

from mygenerativerecommenders_b200 import ops
from oracle import reference_port as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_reference_known_answers(golden):
    g = golden("ops")
    off = ops.asynchronous_complete_cumsum(g["kat_cumsum_in"].to(DEV))
    assert off.dtype == torch.int32 and torch.equal(off.cpu(), g["kat_cumsum_out"])
    assert torch.equal(ops.dense_to_jagged(g["kat_d2j_in"].to(DEV), off).cpu(), g["kat_d2j_out"])
    out = ops.jagged_to_padded_dense(g["kat_j2d_in"].to(DEV), g["kat_j2d_off"].to(DEV), 3, 0)
    assert torch.equal(out.cpu(), g["kat_j2d_out"])


def test_ragged_with_empty_sequences(golden):
    g = golden("ops")
    off = ops.asynchronous_complete_cumsum(g["r_lengths"].to(DEV))
    assert torch.equal(off.cpu(), g["r_offsets"])
    jag = ops.dense_to_jagged(g["r_dense"].to(DEV), off)
    assert torch.equal(jag.cpu(), g["r_jagged"])
    assert torch.equal(ops.jagged_to_padded_dense(jag, off, 12, 0.0).cpu(), g["r_padded"])
    assert torch.equal(ops.jagged_to_padded_dense(jag, off, 12, -1.5).cpu(), g["r_padded_pad"])
    cur = ops.get_current_embeddings(g["cur_lengths"].to(DEV), g["r_dense"].to(DEV))
    assert torch.equal(cur.cpu(), g["cur_out"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.int64, torch.bool, torch.float64])
@pytest.mark.parametrize("width", [1, 3, 50, 64])
@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
def test_conversions_every_dtype_and_alignment(dtype, width, idx):
    gen = torch.Generator().manual_seed(width)
    B, N = 7, 19
    lengths = torch.randint(0, N + 1, (B,), generator=gen).to(idx)
    dense = (torch.randn(B, N, width, generator=gen) * 10)
    dense = (dense > 0) if dtype == torch.bool else dense.to(dtype)
    off_ref = O.complete_cumsum(lengths)
    off = ops.asynchronous_complete_cumsum(lengths.to(DEV))
    assert off.dtype == idx and torch.equal(off.cpu(), off_ref)
    jag = ops.dense_to_jagged(dense.to(DEV), off)
    assert torch.equal(jag.cpu(), O.dense_to_jagged(dense, off_ref))
    back = ops.jagged_to_padded_dense(jag, off, N, 0.0)
    assert back.dtype == dtype
    assert torch.equal(back.cpu(), O.jagged_to_padded_dense(jag.cpu(), off_ref, N, 0.0))


def test_two_dimensional_and_sliced_inputs():
    gen = torch.Generator().manual_seed(5)
    B, N, D = 6, 13, 10
    lengths = torch.randint(1, N, (B,), generator=gen)
    off_ref = O.complete_cumsum(lengths)
    off = off_ref.to(DEV)
    mask = torch.rand(B, N, generator=gen) > 0.5             # (B, N) bool, ops.py:248
    assert torch.equal(ops.dense_to_jagged(mask.to(DEV), off).cpu(), O.dense_to_jagged(mask, off_ref))
    x = torch.randn(B, N + 1, D, generator=gen)
    xg = x.to(DEV)
    # the slices generative_recommenders.py:409-424 passes: [:, :-1] and [:, 1:]
    for sl in (slice(0, N), slice(1, N + 1)):
        got = ops.dense_to_jagged(xg[:, sl, :], off)
        assert torch.equal(got.cpu(), O.dense_to_jagged(x[:, sl, :], off_ref))


def test_mask_dense_by_aux_mask_reference_cases():
    # /root/reference/tests/test_ops.py:56-139, the four cases
    cases = [
        ([[[1, 1], [2, 2], [3, 3], [4, 4]], [[5, 5], [6, 6], [7, 7], [8, 8]]],
         [[False, True, False, True], [True, False, True, False]], [4, 4],
         [[[2, 2], [4, 4], [0, 0], [0, 0]], [[5, 5], [7, 7], [0, 0], [0, 0]]], [2, 2]),
        ([[[1, 1], [2, 2], [3, 3], [4, 4]], [[5, 5], [6, 6], [0, 0], [0, 0]]],
         [[False, True, False, True], [True, False, False, False]], [4, 2],
         [[[2, 2], [4, 4], [0, 0], [0, 0]], [[5, 5], [0, 0], [0, 0], [0, 0]]], [2, 1]),
        ([[[1, 1], [2, 2]], [[3, 3], [4, 4]]], [[False, False], [False, False]], [2, 2],
         [[[0, 0], [0, 0]], [[0, 0], [0, 0]]], [0, 0]),
        ([[[1, 1], [2, 2]], [[3, 3], [4, 4]]], [[True, True], [True, True]], [2, 2],
         [[[1, 1], [2, 2]], [[3, 3], [4, 4]]], [2, 2]),
    ]
    for dense, mask, lengths, exp, exp_len in cases:
        d = torch.tensor(dense, dtype=torch.float, device=DEV)
        out, nl = ops.mask_dense_by_aux_mask(d, torch.tensor(mask, device=DEV),
                                             torch.tensor(lengths, device=DEV), d.shape[1])
        assert torch.equal(out.cpu(), torch.tensor(exp, dtype=torch.float))
        assert torch.equal(nl.cpu(), torch.tensor(exp_len))


def test_autograd_is_the_transposed_copy():
    gen = torch.Generator().manual_seed(9)
    B, N, D = 5, 11, 8
    lengths = torch.tensor([3, 0, 11, 7, 1])
    off = O.complete_cumsum(lengths).to(DEV)
    x = torch.randn(B, N, D, generator=gen).to(DEV).requires_grad_(True)
    w = torch.randn(int(lengths.sum()), D, generator=gen).to(DEV)
    (ops.dense_to_jagged(x, off) * w).sum().backward()
    exp = O.jagged_to_padded_dense(w.cpu(), off.cpu(), N, 0.0)
    assert torch.equal(x.grad.cpu(), exp)
    v = torch.randn(int(lengths.sum()), D, generator=gen).to(DEV).requires_grad_(True)
    w2 = torch.randn(B, N, D, generator=gen).to(DEV)
    (ops.jagged_to_padded_dense(v, off, N, 0.0) * w2).sum().backward()
    assert torch.equal(v.grad.cpu(), O.dense_to_jagged(w2.cpu(), off.cpu()))
    enc = torch.randn(B, N, D, generator=gen).to(DEV).requires_grad_(True)
    ln = torch.tensor([3, 2, 11, 7, 1], device=DEV)
    ops.get_current_embeddings(ln, enc).sum().backward()
    exp = torch.zeros(B, N, D)
    exp[torch.arange(B), ln.cpu() - 1] = 1
    assert torch.equal(enc.grad.cpu(), exp)


def test_full_size_round_trip_properties():
    # C5-shaped: B=128, N=8192, 512 bf16 columns (1 GiB dense); property checks only
    B, N, W = 128, 8192, 512
    gen = torch.Generator().manual_seed(0)
    lengths = torch.randint(1024, N + 1, (B,), generator=gen)
    off = ops.asynchronous_complete_cumsum(lengths.to(DEV))
    assert int(off[-1]) == int(lengths.sum()) and int(off[0]) == 0
    assert torch.equal(off[1:] - off[:-1], lengths.to(DEV))
    T = int(off[-1])
    jag = torch.randn(T, W, device=DEV, dtype=torch.bfloat16)
    dense = ops.jagged_to_padded_dense(jag, off, N, 0.0)
    # (1) jagged -> dense -> jagged is the identity
    assert torch.equal(ops.dense_to_jagged(dense, off, total=T), jag)
    # (2) padding rows are exactly zero and nothing else changed: checksum of checksums
    def csum(t):  # exact integer checksum of the raw bf16 bit patterns
        return t.contiguous().view(torch.int16).long().sum(dim=tuple(range(1, t.dim()))).sum()
    assert int(csum(dense)) == int(csum(jag))
    valid = torch.arange(N, device=DEV).unsqueeze(0) < lengths.to(DEV).unsqueeze(1)
    assert int((dense[~valid] != 0).sum()) == 0
    # (3) idempotence of the composite
    again = ops.jagged_to_padded_dense(ops.dense_to_jagged(dense, off, total=T), off, N, 0.0)
    assert torch.equal(again, dense)


def test_large_batch_cumsum():
    lengths = torch.randint(0, 1000, (100_003,), dtype=torch.int64)
    off = ops.asynchronous_complete_cumsum(lengths.to(DEV))
    assert torch.equal(off.cpu(), O.complete_cumsum(lengths))


@pytest.mark.parametrize("D", [256, 50, 128])
def test_embedding_lookup_gradient_matches_nn_embedding(D):
    """functional.embedding_lookup == F.embedding incl. padding_idx and repeated ids
    (models/embeddings/embeddings.py:40-101 tables); the gradient is a scatter-add."""
    from mygenerativerecommenders_b200 import functional as GF
    g = torch.Generator(device="cuda").manual_seed(D)
    V = 1000
    w = torch.randn(V, D, device="cuda", generator=g)
    ids = torch.randint(0, 40, (7, 33), device="cuda", generator=g)   # many repeats, some zeros
    go = torch.randn(7, 33, D, device="cuda", generator=g)
    w1 = w.clone().requires_grad_(True)
    out1 = GF.embedding_lookup(w1, ids, 0)
    out1.backward(go)
    w2 = w.clone().requires_grad_(True)
    out2 = torch.nn.functional.embedding(ids, w2, padding_idx=0)
    out2.backward(go)
    assert torch.equal(out1, out2)
    assert w1.grad[0].abs().max() == 0
    torch.testing.assert_close(w1.grad, w2.grad, rtol=1e-5, atol=1e-5)


def test_inbatch_prefix_cache_equals_boolean_mask_path():
    """InBatchNegativesSampler.process_batch_prefix (no nonzero / no sync) builds the same cache as
    the reference-shaped process_batch (negative_sampler.py:160-190)."""
    from mygenerativerecommenders_b200 import ops
    from mygenerativerecommenders_b200.negative_sampler import InBatchNegativesSampler
    g = torch.Generator().manual_seed(0)
    B, N, D = 9, 17, 24
    lengths = torch.randint(0, N - 1, (B,), generator=g)
    ids = torch.zeros(B, N, dtype=torch.int64)
    for b in range(B):
        ids[b, : lengths[b] + 1] = torch.randint(1, 30, (int(lengths[b]) + 1,), generator=g)
    ids = ids.cuda()
    table = torch.randn(31, D, generator=g).cuda()
    emb = table[ids]
    for dedup in (False, True):
        a = InBatchNegativesSampler(True, 1e-6, dedup)
        b_ = InBatchNegativesSampler(True, 1e-6, dedup)
        flat = ids.view(-1)
        a.process_batch(flat, flat != 0, emb.view(-1, D))
        off = ops.asynchronous_complete_cumsum(lengths.cuda() + 1)
        b_.process_batch_prefix(ids, emb, off, int(lengths.sum()) + B)
        assert torch.equal(a._cached_ids, b_._cached_ids)
        assert torch.equal(a._cached_embeddings, b_._cached_embeddings)


@pytest.mark.parametrize("W", [256, 50])
def test_l2_normalize_matches_reference_composite(W):
    """functional.l2_normalize == x / clamp(norm(x), min=eps) (negative_sampler.py:31-37), forward and
    backward, including rows below the clamp and a strided input."""
    from mygenerativerecommenders_b200 import functional as GF
    g = torch.Generator(device="cuda").manual_seed(W)
    base = torch.randn(37, W + 6, device="cuda", generator=g)
    base[3] = 0.0
    base[5] *= 1e-9                      # norm below eps: the clamp branch
    go = torch.randn(37, W, device="cuda", generator=g)
    a = base.clone().requires_grad_(True)
    b = base.clone().requires_grad_(True)
    ya = GF.l2_normalize(a[:, 3:3 + W], 1e-6)
    xb = b[:, 3:3 + W]
    yb = xb / torch.clamp(torch.linalg.norm(xb, ord=2, dim=-1, keepdim=True), min=1e-6)
    torch.testing.assert_close(ya, yb, rtol=1e-6, atol=1e-7)
    ya.backward(go)
    yb.backward(go)
    torch.testing.assert_close(a.grad, b.grad, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64])
def test_complete_cumsum_across_tile_boundaries(dtype):
    """a1 at sizes around the kernel's tiles (16 lengths per thread, 16 384 per tile, running carry)."""
    gen = torch.Generator().manual_seed(1)
    for n in (0, 1, 15, 16, 17, 1023, 1024, 16383, 16384, 16385, 50_001):
        x = torch.randint(0, 300, (n,), generator=gen).to(dtype)
        want = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(x.long(), 0)]).to(dtype)
        got = ops.asynchronous_complete_cumsum(x.to(DEV))
        assert got.dtype == dtype and torch.equal(got.cpu(), want), n
