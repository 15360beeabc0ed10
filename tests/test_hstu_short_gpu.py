"""Short-sequence attention kernels (csrc/hstu_attn_short.cu: n <= 256, bf16, 64-wide heads) against
the fp64 padded oracle (hstu.py:96-128, :134-205) and against the long-sequence tcgen05 kernels on
the same inputs.  Tolerances: DESIGN.md §2 (bf16 path: 1e-2 max / 5e-3 rel-L2 outputs, 2e-2 grads)."""
import pytest
import torch

from mygenerativerecommenders_b200 import functional as GF
from mygenerativerecommenders_b200 import hstu
from oracle import reference_port as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
D = 64


def _close(got, ref, tol_inf, tol_l2, what):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    scale = max(ref.abs().max().item(), 1e-12)
    err = (got - ref).abs().max().item()
    assert err <= tol_inf * scale, f"{what}: max|d|={err:.3e} > {tol_inf:.0e}*{scale:.3e}"
    rel = ((got - ref).norm() / max(ref.norm().item(), 1e-12)).item()
    assert rel <= tol_l2, f"{what}: rel-l2 {rel:.3e} > {tol_l2:.0e}"


def _case(seed, N, H, lengths, with_ts=True, ts_step=5000):
    gen = torch.Generator().manual_seed(seed)
    B = len(lengths)
    lengths = torch.tensor(lengths, dtype=torch.int64)
    off = O.complete_cumsum(lengths)
    T = int(off[-1])
    q, k, v = ((torch.randn(T, H * D, generator=gen) * 0.5).to(torch.bfloat16).float() for _ in range(3))
    ts = None
    if with_ts:
        ts = 978_300_000 + torch.cumsum(torch.randint(1, ts_step, (B, N), generator=gen), dim=1)
        ts = ts * (torch.arange(N).unsqueeze(0) <= lengths.unsqueeze(1))
    ts_w = torch.randn(129, generator=gen) * 0.5
    pos_w = torch.randn(2 * N - 1, generator=gen) * 0.5
    w = torch.randn(T, H * D, generator=gen).to(torch.bfloat16).float()
    return dict(off=off, T=T, q=q, k=k, v=v, ts=ts, ts_w=ts_w, pos_w=pos_w, w=w, lengths=lengths)


_THR = None


def _thr():
    global _THR
    if _THR is None:
        _THR = hstu.tabulate_bucket_thresholds(hstu._default_bucketization, 128).to(DEV)
    return _THR


def _run(c, N, H, grad=True):
    q, k, v = (c[n].to(DEV).to(torch.bfloat16).requires_grad_(grad) for n in ("q", "k", "v"))
    with_ts = c["ts"] is not None
    ts_w = c["ts_w"].to(DEV).requires_grad_(grad)
    pos_w = c["pos_w"].to(DEV).requires_grad_(grad)
    out = GF.hstu_attention(q, k, v, c["off"].to(DEV), c["ts"].to(DEV) if with_ts else None,
                            ts_w if with_ts else None, pos_w if with_ts else None,
                            _thr() if with_ts else None, N, H, D, D)
    if grad:
        out.backward(c["w"].to(DEV).to(torch.bfloat16))
    return out, (q, k, v, ts_w, pos_w)


def _oracle(c, N, H):
    with_ts = c["ts"] is not None
    leaves = [c[n].clone().double().requires_grad_(True) for n in ("q", "k", "v", "ts_w", "pos_w")]
    ref = O.hstu_attention(leaves[0], leaves[1], leaves[2], c["off"], c["ts"],
                           leaves[3] if with_ts else None, leaves[4] if with_ts else None, N, H, D, D)
    (ref * c["w"].double()).sum().backward()
    return ref, leaves


CASES = [
    # N, H, lengths, with_ts
    (211, 4, [211, 37, 129, 128, 1, 0, 64, 65, 127, 200], True),    # C2 shape, tile / half-tile edges
    (256, 2, [256, 255, 130, 3], True),                             # the largest the path takes
    (61, 1, [61, 5, 50, 33, 17], True),                             # C3 shape: one tile, one head
    (128, 2, [128, 100, 64], True),                                 # single-phase grid (max_len <= 128)
    (211, 2, [190, 77, 140], False),                                # no timestamps: masks only
    (100, 8, [100, 31, 32, 33, 96, 97], True),
]


@pytest.mark.parametrize("N,H,lengths,with_ts", CASES)
def test_short_forward_backward_vs_fp64_oracle(N, H, lengths, with_ts):
    assert GF.short_path_applies(torch.empty(0, dtype=torch.bfloat16), D, D, N)
    c = _case(N * 7 + H, N, H, lengths, with_ts)
    out, leaves = _run(c, N, H)
    ref, rl = _oracle(c, N, H)
    _close(out, ref, 1e-2, 5e-3, "short fwd")
    names = ("dq", "dk", "dv") + (("d_ts_w", "d_pos_w") if with_ts else ())
    for name, got, r in zip(names, leaves, rl):
        _close(got.grad, r.grad, 2e-2, 2e-2, f"short {name}")


@pytest.mark.parametrize("N,H,lengths", [(211, 4, [211, 37, 129, 128, 1, 64, 200]), (256, 2, [256, 130])])
def test_short_kernels_agree_with_the_long_sequence_kernels(N, H, lengths, monkeypatch):
    c = _case(99 + N, N, H, lengths)
    out_s, leaves_s = _run(c, N, H)
    monkeypatch.setenv("GRB_NO_SHORT", "1")
    out_l, leaves_l = _run(c, N, H)
    monkeypatch.delenv("GRB_NO_SHORT")
    # same bf16 inputs, same fp16 SiLU in the forward; the bias is rounded to fp16 here
    _close(out_s, out_l, 8e-3, 2e-3, "fwd short vs long")
    for name, a, b_ in zip(("dq", "dk", "dv"), leaves_s, leaves_l):
        _close(a.grad, b_.grad, 1e-2, 5e-3, f"{name} short vs long")
    # bias gradients: the short path sums the heads' bf16 dS in bf16 before binning, the long path
    # bins bf16 dS per head in fp32; both are checked against the fp64 oracle at 2e-2 above
    for name, a, b_ in zip(("d_ts_w", "d_pos_w"), leaves_s[3:], leaves_l[3:]):
        _close(a.grad, b_.grad, 2e-2, 1e-2, f"{name} short vs long")


def test_short_kernels_are_deterministic_in_q_k_v_and_stable_under_repeats():
    """Race hunting without a sanitizer: dq/dk/dv/out have no atomics, so repeated launches over
    ragged random batches must be bit-identical; the bias gradients (heads summed in bf16 by bulk
    reduce-adds in arrival order, then fp32 atomics) within 5e-3."""
    gen = torch.Generator().manual_seed(5)
    for trial in range(10):
        H = [1, 2, 4][trial % 3]
        N = int(torch.randint(20, 257, (1,), generator=gen))
        B = int(torch.randint(1, 40, (1,), generator=gen))
        lengths = torch.randint(0, N + 1, (B,), generator=gen).tolist()
        lengths[0] = N
        c = _case(300 + trial, N, H, lengths)
        runs = []
        for _ in range(3):
            out, leaves = _run(c, N, H)
            runs.append([out.detach()] + [l.grad.clone() for l in leaves])
        for other in runs[1:]:
            for idx in range(4):
                assert torch.equal(runs[0][idx], other[idx]), (trial, idx)
            for idx in (4, 5):
                _close(other[idx], runs[0][idx], 5e-3, 5e-3, "bias grads repeat")
        assert all(torch.isfinite(t).all() for t in runs[0])


def test_short_path_full_c2_shape_batch():
    """128 sequences x U[20, 200] tokens, 4 heads: the launch bench.py times (both grid phases,
    two CTAs per SM, ~300 CTAs in flight)."""
    gen = torch.Generator().manual_seed(11)
    N, H = 211, 4
    lengths = torch.randint(20, 201, (128,), generator=gen).tolist()
    c = _case(1234, N, H, lengths)
    out, leaves = _run(c, N, H)
    ref, rl = _oracle(c, N, H)
    _close(out, ref, 1e-2, 5e-3, "C2 fwd")
    for name, got, r in zip(("dq", "dk", "dv", "d_ts_w", "d_pos_w"), leaves, rl):
        _close(got.grad, r.grad, 2e-2, 2e-2, f"C2 {name}")


def test_masked_bucket_tiles_and_item_schedule():
    """grb_hstu_bucket_tiles_masked: the reference's bucket (hstu.py:113-123) for j <= i < n, 255
    elsewhere, both orientations; grb_hstu_short_schedule: non-empty sequences, n > 128 first."""
    N, lengths = 211, [211, 140, 128, 5, 0]
    c = _case(3, N, 1, lengths)
    off, ts = c["off"].to(DEV), c["ts"].to(DEV)
    cache = GF.hstu_bucket_cache(off, ts, _thr(), N, masked=True)
    sched = cache.grb_sched.cpu()
    assert sched[0] == 4 and sched[1] == 2
    assert sorted(sched[2:4].tolist()) == [0, 1] and sorted(sched[4:6].tolist()) == [2, 3]
    t = cache.cpu().view(len(lengths), 3, 2, 8, 128, 16)
    ext = torch.cat([c["ts"], c["ts"][:, N - 1:N]], dim=1)
    ref = O.bucketize_ts(ext[:, 1:].unsqueeze(2) - ext[:, :-1].unsqueeze(1))     # (B, N, N)
    for b, n in enumerate(lengths):
        for slot, (iq, jk) in enumerate([(0, 0), (1, 0), (1, 1)]):
            if iq * 128 >= n:
                continue
            tq = t[b, slot, 0].permute(1, 0, 2).reshape(128, 128)          # [query row][key col]
            tk = t[b, slot, 1].permute(1, 0, 2).reshape(128, 128).t()      # stored [key][query]
            i = iq * 128 + torch.arange(128).view(-1, 1)
            j = jk * 128 + torch.arange(128).view(1, -1)
            valid = (j <= i) & (i < n)
            want = torch.full((128, 128), 255, dtype=torch.uint8)
            ii, jj = i.expand(128, 128)[valid], j.expand(128, 128)[valid]
            want[valid] = ref[b, ii, jj].to(torch.uint8)
            assert torch.equal(tq, want) and torch.equal(tk, want), (b, slot)
    # mask-only tiles (no timestamps): valid pairs hold 0
    only = GF.hstu_bucket_cache(off, None, None, N, masked=True).cpu().view(len(lengths), 3, 2, 8, 128, 16)
    tq = only[1, 1, 0].permute(1, 0, 2).reshape(128, 128)                  # sequence of 140, slot (1, 0)
    assert (tq[:12] == 0).all() and (tq[12:] == 255).all()


@pytest.mark.parametrize("with_ts", [True, False])
def test_row_bucket_tail_is_zeroed_by_the_kernels_themselves(with_ts):
    """Fixed-size row buckets (rows_padded): q / k / v carry rows past offsets[-1]; the short-sequence
    launches zero those rows of out / dq / dk / dv themselves (grb_hstu_attn_args.zero_tail_rows) — no
    separate fill.  The allocator is poisoned with NaN first so that an unwritten row shows."""
    N, H, lengths = 211, 4, [200, 37, 129, 5]
    c = _case(21, N, H, lengths, with_ts=with_ts)
    T, pad = c["T"], 300
    ref_out, ref_leaves = _run(c, N, H)
    for _ in range(4):                                   # poison blocks of the sizes about to be allocated
        for rows in (T + pad, 3 * (T + pad)):
            torch.full((rows, H * D), float("nan"), dtype=torch.bfloat16, device=DEV)
    q, k, v = (torch.cat([c[n], torch.zeros(pad, H * D)]).to(DEV).to(torch.bfloat16).requires_grad_(True)
               for n in ("q", "k", "v"))
    ts_w = c["ts_w"].to(DEV).requires_grad_(True)
    pos_w = c["pos_w"].to(DEV).requires_grad_(True)
    out = GF.hstu_attention(q, k, v, c["off"].to(DEV), c["ts"].to(DEV) if with_ts else None,
                            ts_w if with_ts else None, pos_w if with_ts else None,
                            _thr() if with_ts else None, N, H, D, D, rows_padded=True)
    assert out.shape[0] == T + pad and not out[T:].any() and torch.isfinite(out).all()
    assert torch.equal(out[:T], ref_out)
    out.backward(torch.cat([c["w"], torch.randn(pad, H * D)]).to(DEV).to(torch.bfloat16))
    for g, r in zip((q, k, v), ref_leaves[:3]):
        assert not g.grad[T:].any() and torch.equal(g.grad[:T], r.grad)
