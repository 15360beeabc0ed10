"""f2: the jagged input path (csrc/input_path.cu) against the reference's padded composite
(embeddings.py:94-97, learnable_positional_embedding.py:42-58, hstu.py:502, postprocessors.py:47-55)
restated on torch ops.  fp32 outputs: exact up to one fma rounding (1e-6 relative); bf16: one bf16
rounding (4e-3); gradients: fp32 atomics reorder sums (1e-5 relative to the largest entry)."""
import pytest
import torch

from mygenerativerecommenders_b200 import functional as GF
from mygenerativerecommenders_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(B=9, N=37, D=64, V=200, seed=0, lengths=None):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.tensor(lengths) if lengths is not None else torch.randint(0, N, (B,), generator=g)
    ids = torch.randint(1, V, (B, N), generator=g) * (torch.arange(N).unsqueeze(0) <= lengths.unsqueeze(1))
    table = torch.randn(V, D, generator=g)
    table[0] = 0
    pos = torch.randn(N, D, generator=g)
    return lengths.to(DEV), ids.to(DEV), table.to(DEV), pos.to(DEV)


def _reference(table, pos, ids, lengths, scale):
    """The reference's padded formulation, then dense_to_jagged."""
    x = torch.nn.functional.embedding(ids, table) * scale + pos[: ids.shape[1]].unsqueeze(0)
    x = x * (ids != 0).unsqueeze(-1).float()
    rows = [x[b, : int(lengths[b])] for b in range(ids.shape[0])]
    return torch.cat(rows, 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("extra_rows", [0, 19])
def test_forward_and_backward_without_dropout(dtype, extra_rows):
    lengths, ids, table, pos = _setup(lengths=[36, 0, 5, 17, 1, 36, 20, 9, 3])
    off = ops.asynchronous_complete_cumsum(lengths)
    T = int(lengths.sum())
    scale = 64 ** 0.5
    t1, p1 = table.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    y = GF.jagged_input(t1, p1, ids, off, T + extra_rows, scale, 0.0, None, out_dtype=dtype)
    t2, p2 = table.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    ref = _reference(t2, p2, ids, lengths, scale)
    assert y.shape == (T + extra_rows, 64) and y.dtype == dtype
    tol = 1e-6 if dtype == torch.float32 else 4e-3
    assert (y[:T].float() - ref).abs().max().item() <= tol * ref.abs().max().item()
    assert (y[T:] == 0).all()
    w = torch.randn(T + extra_rows, 64, device=DEV).to(dtype)
    y.backward(w)
    ref.backward(w[:T].float())
    for got, want in ((t1.grad, t2.grad), (p1.grad, p2.grad)):
        assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    assert t1.grad[0].abs().max() == 0            # id 0 = padding row: no gradient


def test_dropout_mask_statistics_scaling_and_backward_consistency():
    lengths, ids, table, pos = _setup(B=64, N=100, D=256, V=3000, seed=1)
    off = ops.asynchronous_complete_cumsum(lengths)
    T = int(lengths.sum())
    seed = torch.tensor([12345], dtype=torch.int64, device=DEV)
    t1 = table.clone().requires_grad_(True)
    y0 = GF.jagged_input(table, pos, ids, off, T, 16.0, 0.0, None)
    y = GF.jagged_input(t1, pos, ids, off, T, 16.0, 0.2, seed)
    kept = y != 0
    frac = kept.float().mean().item()
    assert abs(frac - 0.8) < 5e-3, frac                                   # 1.6 M elements: sigma = 3e-4
    assert torch.allclose(y[kept], y0[kept] / 0.8, rtol=1e-6, atol=0)     # survivors scaled by 1 / (1 - p)
    # rows and columns are hit evenly (no structure from the counter layout)
    assert (kept.float().mean(0) - 0.8).abs().max() < 0.03 and (kept.float().mean(1) - 0.8).abs().max() < 0.12
    # same seed -> same mask; another seed -> another mask
    y_again = GF.jagged_input(table, pos, ids, off, T, 16.0, 0.2, seed)
    y_other = GF.jagged_input(table, pos, ids, off, T, 16.0, 0.2, seed + 1)
    assert torch.equal(y, y_again) and not torch.equal(y, y_other)
    # the backward regenerates the same mask: d(table) is the scatter-add of mask * scale / (1 - p)
    y.backward(torch.ones_like(y))
    flat = torch.cat([ids[b, : int(lengths[b])] for b in range(ids.shape[0])])
    want = torch.zeros_like(table).index_add_(0, flat, kept.float() * (16.0 / 0.8))
    assert torch.allclose(t1.grad, want, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_l2norm_postprocessor_takes_the_compute_dtype(dtype):
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(300, 256, generator=g) * 3).to(dtype).to(DEV)
    x[5] = 0                                                              # clamp branch
    x1 = x.clone().requires_grad_(True)
    y = GF.l2_normalize(x1, 1e-6)
    assert y.dtype == torch.float32
    x2 = x.float().clone().requires_grad_(True)
    ref = x2 / torch.clamp(torch.linalg.norm(x2, dim=-1, keepdim=True), min=1e-6)
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-6)
    w = torch.randn(300, 256, generator=g).to(DEV)
    y.backward(w)
    ref.backward(w)
    assert x1.grad.dtype == dtype
    tol = 1e-5 if dtype == torch.float32 else 8e-3
    assert (x1.grad.float() - x2.grad).abs().max().item() <= tol * x2.grad.abs().max().item()


def test_table_grad_scope_equals_autograd_accumulation():
    """GF.TableGradScope: three readers of one table scatter into ONE dense buffer; the gradient equals
    what autograd's add of three dense gradients gives, an unused reader is skipped, and a table that
    needs no gradient takes the plain path."""
    lengths, ids, table, pos = _setup(lengths=[36, 0, 5, 17, 1, 36, 20, 9, 3])
    off = ops.asynchronous_complete_cumsum(lengths)
    T = int(lengths.sum())
    pick = torch.randint(0, 200, (77,), device=DEV)
    pick2 = torch.randint(1, 200, (5, 11), device=DEV)
    w = [torch.randn(T, 64, device=DEV), torch.randn(77, 64, device=DEV), torch.randn(5, 11, 64, device=DEV)]

    def run(use_scope, use_third=True):
        t = table.clone().requires_grad_(True)
        p = pos.clone().requires_grad_(True)
        sc = GF.TableGradScope(t) if use_scope else None
        a = GF.jagged_input(t, p, ids, off, T, 8.0, 0.0, None, grad_scope=sc)
        b = GF.embedding_lookup(t, pick, 0, grad_scope=sc)
        c = GF.embedding_lookup(t, pick2, 0, grad_scope=sc)
        loss = (a * w[0]).sum() + (b * w[1]).sum()
        if use_third:
            loss = loss + (c * w[2]).sum()
        loss.backward()
        return t.grad, p.grad

    for third in (True, False):
        (g0, p0), (g1, p1) = run(False, third), run(True, third)
        assert torch.allclose(g0, g1, rtol=1e-5, atol=1e-5) and torch.allclose(p0, p1, rtol=1e-5, atol=1e-5)
    with torch.no_grad():
        sc = GF.TableGradScope(table)
        assert sc.proxy is None
        assert torch.equal(GF.embedding_lookup(table, pick, 0, grad_scope=sc), table[pick])


def test_zero_tail_rows_and_fused_negative_draw():
    """grb_zero_tail_rows clears exactly the rows past offsets[-1]; grb_draw_negatives draws uniformly over
    [0, count) with the count read on the device and gathers the cached ids in the same pass."""
    import ctypes as C
    from mygenerativerecommenders_b200 import _lib
    from mygenerativerecommenders_b200.negative_sampler import InBatchNegativesSampler
    off = torch.tensor([0, 5, 5, 37], dtype=torch.int64, device=DEV)
    t = torch.full((3, 64, 24), 7.0, dtype=torch.bfloat16, device=DEV)
    GF._zero_tail_rows(t, 3, off)
    assert (t[:, :37] == 7).all() and (t[:, 37:] == 0).all()
    smp = InBatchNegativesSampler(l2_norm=True, l2_norm_eps=1e-6, dedup_embeddings=True)
    cached = torch.arange(1000, 1000 + 4096, device=DEV)
    smp._cached_ids, smp._cached_count = cached, torch.tensor(777, device=DEV)
    smp._cached_embeddings = torch.zeros(4096, 8, device=DEV)
    pos = torch.zeros(3001, dtype=torch.int64, device=DEV)
    offs, ids = smp._draw_with_ids(pos, 129)
    assert offs.shape == (3001, 129) and offs.min() >= 0 and offs.max() == 776
    assert torch.equal(ids, cached[offs])
    hist = torch.bincount(offs.view(-1), minlength=777).float()
    exp = offs.numel() / 777
    assert ((hist - exp).abs() < 6 * exp ** 0.5).all()              # ~498 per bin, sigma 22
    offs2, _ = smp._draw_with_ids(pos, 129)
    assert not torch.equal(offs, offs2)                             # a fresh seed per call
    smp._draw = lambda p, n: torch.zeros(p.size(0), n, dtype=torch.int64, device=DEV)   # injected draws win
    assert smp._draw_with_ids(pos, 4)[0].abs().sum() == 0
