"""a4-a8 on the GPU: fused jagged attention, LN/gate and the HSTU modules against the reference's
golden outputs (tests/golden/hstu.pt) and the CPU oracle.

Tolerances: fp32 path  max|d| <= 1e-5 * ||ref||_inf for outputs, 2e-4 for gradients
(fp32 atomics reorder sums); bf16 path  ||d||_inf <= 1e-2 * ||ref||_inf and ||d||_2/||ref||_2 <=
5e-3 for outputs, 2e-2 for gradients."""
import pytest
import torch

from mygenerativerecommenders_b200 import functional as GF
from mygenerativerecommenders_b200 import hstu
from oracle import reference_port as O
from conftest import hstu_case, hstu_incremental_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, ref, tol_inf, tol_l2=None, what=""):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    scale = max(ref.abs().max().item(), 1e-12)
    err = (got - ref).abs().max().item()
    assert err <= tol_inf * scale, f"{what}: max|d|={err:.3e} > {tol_inf:.0e}*{scale:.3e}"
    if tol_l2 is not None:
        rel = ((got - ref).norm() / max(ref.norm().item(), 1e-12)).item()
        assert rel <= tol_l2, f"{what}: rel-l2 {rel:.3e} > {tol_l2:.0e}"


def _rand_case(seed, B, N, H, dqk, dv, lengths, with_ts=True, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    lengths = torch.tensor(lengths, dtype=torch.int64)
    off = O.complete_cumsum(lengths)
    T = int(off[-1])
    q = torch.randn(T, H * dqk, generator=gen)
    k = torch.randn(T, H * dqk, generator=gen)
    v = torch.randn(T, H * dv, generator=gen)
    ts = None
    if with_ts:
        steps = torch.randint(1, 100000, (B, N), generator=gen)
        ts = 978_300_000 + torch.cumsum(steps, dim=1)
        mask = torch.arange(N).unsqueeze(0) <= lengths.unsqueeze(1)  # keep index n_b (target ts)
        ts = ts * mask
    ts_w = torch.randn(129, generator=gen) * 0.5
    pos_w = torch.randn(2 * N - 1, generator=gen) * 0.5
    return dict(off=off, T=T, q=q, k=k, v=v, ts=ts, ts_w=ts_w, pos_w=pos_w, lengths=lengths)


THR = None


def _thr():
    global THR
    if THR is None:
        THR = hstu.tabulate_bucket_thresholds(hstu._default_bucketization, 128).to(DEV)
    return THR


def _run_kernel(c, N, H, dqk, dv, dtype=torch.float32, grad=False):
    q, k, v = (c[n].to(DEV).to(dtype).requires_grad_(grad) for n in ("q", "k", "v"))
    ts = c["ts"].to(DEV) if c["ts"] is not None else None
    ts_w = c["ts_w"].to(DEV).requires_grad_(grad)
    pos_w = c["pos_w"].to(DEV).requires_grad_(grad)
    out = GF.hstu_attention(q, k, v, c["off"].to(DEV), ts, ts_w if ts is not None else None,
                            pos_w if ts is not None else None, _thr() if ts is not None else None,
                            N, H, dqk, dv)
    return out, (q, k, v, ts_w, pos_w)


@pytest.mark.parametrize("B,N,H,dqk,dv,lengths", [
    (4, 24, 2, 8, 8, [1, 24, 7, 13]),
    (3, 211, 1, 50, 50, [211, 37, 64]),           # ml-1m head shape, N not a tile multiple
    (5, 70, 3, 16, 24, [0, 65, 64, 1, 70]),        # empty sequence, tile-edge lengths, dqk != dv
    (2, 130, 2, 64, 64, [130, 129]),
    (2, 40, 1, 200, 130, [40, 17]),                # wide heads (forward only)
])
def test_attention_forward_fp32_vs_oracle(B, N, H, dqk, dv, lengths):
    c = _rand_case(B * 100 + N, B, N, H, dqk, dv, lengths)
    out, _ = _run_kernel(c, N, H, dqk, dv)
    ref = O.hstu_attention(c["q"], c["k"], c["v"], c["off"], c["ts"], c["ts_w"], c["pos_w"],
                           N, H, dqk, dv)
    _close(out, ref, 1e-5, what="attn fwd")


def test_attention_without_timestamps_has_no_bias():
    c = _rand_case(7, 3, 33, 2, 16, 16, [33, 5, 20], with_ts=False)
    out, _ = _run_kernel(c, 33, 2, 16, 16)
    ref = O.hstu_attention(c["q"], c["k"], c["v"], c["off"], None, None, None, 33, 2, 16, 16)
    _close(out, ref, 1e-5, what="attn fwd no-ts")


@pytest.mark.parametrize("B,N,H,dqk,dv,lengths", [
    (4, 24, 2, 8, 8, [1, 24, 7, 13]),
    (3, 211, 1, 50, 50, [211, 37, 64]),
    (4, 70, 3, 16, 24, [0, 65, 64, 70]),
    (2, 130, 2, 64, 64, [130, 129]),
])
def test_attention_backward_fp32_vs_oracle(B, N, H, dqk, dv, lengths):
    c = _rand_case(B * 10 + N, B, N, H, dqk, dv, lengths)
    gen = torch.Generator().manual_seed(1)
    w = torch.randn(c["T"], H * dv, generator=gen)
    out, leaves = _run_kernel(c, N, H, dqk, dv, grad=True)
    (out * w.to(DEV)).sum().backward()
    ref_leaves = [c[n].clone().double().requires_grad_(True) for n in ("q", "k", "v", "ts_w", "pos_w")]
    ref = O.hstu_attention(ref_leaves[0], ref_leaves[1], ref_leaves[2], c["off"], c["ts"],
                           ref_leaves[3], ref_leaves[4], N, H, dqk, dv)
    (ref * w.double()).sum().backward()
    for name, got, r in zip(("dq", "dk", "dv", "d_ts_w", "d_pos_w"), leaves, ref_leaves):
        _close(got.grad, r.grad, 2e-4, what=name)


def test_padded_rows_and_neighbour_sequences_do_not_leak():
    # same sequence alone and inside a batch must give identical rows (jagged independence)
    N, H, d = 96, 2, 32
    c = _rand_case(11, 3, N, H, d, d, [50, 96, 3])
    out, _ = _run_kernel(c, N, H, d, d)
    s, e = int(c["off"][1]), int(c["off"][2])
    solo = dict(c)
    solo.update(off=torch.tensor([0, e - s]), q=c["q"][s:e], k=c["k"][s:e], v=c["v"][s:e],
                ts=c["ts"][1:2], T=e - s)
    out1, _ = _run_kernel(solo, N, H, d, d)
    assert torch.equal(out[s:e], out1)


@pytest.mark.parametrize("W", [50, 64, 256, 512, 1000, 1024])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ln_gate_vs_torch(W, dtype):
    gen = torch.Generator().manual_seed(W)
    x = torch.randn(37, W, generator=gen) * 3 + 1
    u = torch.randn(37, W, generator=gen)
    dy = torch.randn(37, W, generator=gen)
    xg, ug = x.to(DEV).to(dtype).requires_grad_(True), u.to(DEV).to(dtype).requires_grad_(True)
    y = GF.layer_norm_gate(xg, ug, 1e-6)
    y.backward(dy.to(DEV).to(dtype))
    xr, ur = xg.detach().double().cpu().requires_grad_(True), ug.detach().double().cpu().requires_grad_(True)
    yr = ur * torch.nn.functional.layer_norm(xr, [W], eps=1e-6)
    yr.backward(dy.to(dtype).double())
    t = (1e-5, None) if dtype == torch.float32 else (1e-2, 5e-3)
    _close(y, yr, *t, what="ln_gate y")
    tg = (1e-4, None) if dtype == torch.float32 else (2e-2, 2e-2)
    _close(xg.grad, xr.grad, *tg, what="ln_gate dx")
    _close(ug.grad, ur.grad, *tg, what="ln_gate dgate")
    y2 = GF.layer_norm_gate(xg.detach(), None, 1e-6)
    _close(y2, torch.nn.functional.layer_norm(xr.detach(), [W], eps=1e-6), *t, what="ln y")


@pytest.mark.parametrize("W", [256, 512, 1024])
@pytest.mark.parametrize("gated", [True, False])
def test_ln_gate_fused_dropout_mask_is_consistent_forward_and_backward(W, gated):
    """hstu.py:404-408 dropout(u * norm(a)) drawn inside the ln_gate kernels: the kept fraction and the scale
    are those of F.dropout, the same seed gives the same mask, another salt / seed another one, and the
    backward applies exactly the mask the forward drew (checked against torch autograd with the mask read
    off the forward's zeros)."""
    gen = torch.Generator().manual_seed(W)
    rows, p = 301, 0.2
    x = (torch.randn(rows, W, generator=gen) * 2 + 0.5).to(DEV).bfloat16().requires_grad_(True)
    u = (torch.randn(rows, W, generator=gen) + 3.0).to(DEV).bfloat16().requires_grad_(gated)   # u != 0
    dy = torch.randn(rows, W, generator=gen).to(DEV).bfloat16()
    seed = torch.tensor([123456789012345], dtype=torch.int64, device=DEV)
    gate = u if gated else None
    y0 = GF.layer_norm_gate(x, gate, 1e-6)
    y = GF.layer_norm_gate(x, gate, 1e-6, p, seed, 3)
    y.backward(dy)
    keep = y != 0
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 4 * (p * (1 - p) / keep.numel()) ** 0.5 + 2e-3, frac
    # kept entries are the undropped ones scaled by 1 / (1 - p) (bf16 rounding of both sides)
    _close(y[keep].float(), y0[keep].float() / (1 - p), 2e-2, 1e-2, what="kept values")
    assert not keep.all(dim=0).any() and keep.any(dim=1).all()      # no dead column pattern, no dead row
    assert torch.equal(y, GF.layer_norm_gate(x, gate, 1e-6, p, seed, 3))
    assert not torch.equal(y != 0, GF.layer_norm_gate(x, gate, 1e-6, p, seed, 4) != 0)
    assert not torch.equal(y != 0, GF.layer_norm_gate(x, gate, 1e-6, p, seed + 1, 3) != 0)
    # reference gradients with the kernel's own mask
    xr = x.detach().double().requires_grad_(True)
    ur = u.detach().double().requires_grad_(gated)
    yr = torch.nn.functional.layer_norm(xr, [W], eps=1e-6) * (ur if gated else 1.0) * keep.double() / (1 - p)
    yr.backward(dy.double())
    _close(x.grad, xr.grad, 2e-2, 2e-2, what="dropout dx")
    if gated:
        _close(u.grad, ur.grad, 2e-2, 2e-2, what="dropout dgate")


@pytest.mark.parametrize("W,dtype", [(256, torch.bfloat16), (512, torch.bfloat16), (50, torch.float32), (96, torch.bfloat16)])
def test_layer_norm_skip_sums_both_gradients_of_x_in_the_kernel(W, dtype):
    """(LN(x), x) through one autograd node: dx = LN backward + the residual branch's gradient."""
    gen = torch.Generator().manual_seed(W + 1)
    x = torch.randn(77, W, generator=gen).to(DEV).to(dtype).requires_grad_(True)
    g1 = torch.randn(77, W, generator=gen).to(DEV).to(dtype)
    g2 = torch.randn(77, W, generator=gen).to(DEV).to(dtype)
    normed, skip = GF.layer_norm_skip(x, 1e-6)
    assert skip.data_ptr() == x.data_ptr() and torch.equal(skip, x)
    (normed * g1 + skip * g2).sum().backward()
    xr = x.detach().double().requires_grad_(True)
    (torch.nn.functional.layer_norm(xr, [W], eps=1e-6) * g1.double() + xr * g2.double()).sum().backward()
    t = (1e-4, None) if dtype == torch.float32 else (2e-2, 2e-2)
    _close(x.grad, xr.grad, *t, what="ln_skip dx")
    # one branch only
    x2 = x.detach().clone().requires_grad_(True)
    n2, s2 = GF.layer_norm_skip(x2, 1e-6)
    (s2 * g2).sum().backward()
    _close(x2.grad, g2, 1e-6, None, what="skip-only dx")
    x3 = x.detach().clone().requires_grad_(True)
    n3, s3 = GF.layer_norm_skip(x3, 1e-6)
    (n3 * g1).sum().backward()
    x4 = x.detach().clone().requires_grad_(True)
    (GF.layer_norm_gate(x4, None, 1e-6) * g1).sum().backward()
    assert torch.equal(x3.grad, x4.grad)


def _build(c, compute_dtype=None, normalization="rel_bias", linear_activation="silu", **kw):
    enc = hstu.HSTU(max_sequence_len=c["max_seq"], max_output_len=c["out_len"],
                    embedding_dim=c["D"], item_embedding_dim=c["D"], num_blocks=c["blocks"],
                    num_heads=c["H"], linear_dim=c["dv"], attention_dim=c["dqk"],
                    normalization=normalization, linear_config="uvqk", linear_activation=linear_activation, **kw,
                    linear_dropout_rate=0.2, attn_dropout_rate=0.0, compute_dtype=compute_dtype)
    enc.load_state_dict(c["sd"], strict=False)
    return enc.to(DEV).eval()


@pytest.mark.parametrize("name", ["mh", "ml1m", "h64"])
def test_hstu_module_fp32_vs_reference_golden(golden, name):
    c = hstu_case(golden("hstu"), name)
    enc = _build(c)
    x = c["x"].to(DEV).requires_grad_(True)
    y, _ = enc(past_lengths=c["lengths"].to(DEV), user_embeddings=x, valid_mask=None,
               past_payloads={"timestamps": c["ts"].to(DEV)})
    _close(y, c["y"], 1e-5, what=f"{name} y")
    (y * c["w"].to(DEV)).sum().backward()
    _close(x.grad, c["dx"], 2e-4, what=f"{name} dx")
    for k, p in enc.named_parameters():
        _close(p.grad, c["grads"][k], 5e-4, what=f"{name} grad {k}")


def test_hstu_module_bf16_vs_reference_golden(golden):
    c = hstu_case(golden("hstu"), "h64")
    enc = _build(c, compute_dtype=torch.bfloat16)
    x = c["x"].to(DEV).requires_grad_(True)
    y, _ = enc(past_lengths=c["lengths"].to(DEV), user_embeddings=x, valid_mask=None,
               past_payloads={"timestamps": c["ts"].to(DEV)})
    assert y.dtype == torch.float32
    # bf16 outputs 1e-2 (max) / 5e-3 (rel-L2), gradients 2e-2 / 2e-2.  Measured on the
    # B200: y 3.8e-3 / 2.4e-3, dx 3.9e-3 / 2.4e-3, parameter gradients <= 1.3e-2 / 1.0e-2 (_ts_w the largest).
    _close(y, c["y"], 1e-2, 5e-3, what="bf16 y")
    (y * c["w"].to(DEV)).sum().backward()
    _close(x.grad, c["dx"], 2e-2, 2e-2, what="bf16 dx")
    for k, p in enc.named_parameters():
        _close(p.grad, c["grads"][k], 2e-2, 2e-2, what=f"bf16 grad {k}")


def test_shared_weight_copies_and_gradient_buffers_change_nothing(monkeypatch):
    """HSTUJagged._layer_aux (one multi-tensor cast of all master weights, one zero fill for all weight-gradient
    accumulators) against every layer casting / zero-filling for itself: same outputs (bitwise: same operands),
    same gradients (split-K reds: summation order differs, 1e-5)."""
    torch.manual_seed(4)
    B, max_seq, out_len, D, H, d, blocks = 5, 120, 9, 256, 4, 64, 2      # D = 256: the tcgen05 projection GEMMs
    N = max_seq + out_len
    proto = hstu.HSTU(max_sequence_len=max_seq, max_output_len=out_len, embedding_dim=D, item_embedding_dim=D,
                      num_blocks=blocks, num_heads=H, linear_dim=d, attention_dim=d, normalization="rel_bias",
                      linear_config="uvqk", linear_activation="silu", linear_dropout_rate=0.2, attn_dropout_rate=0.0)
    lengths = torch.tensor([129, 40, 77, 3, 110])
    ts = 978_300_000 + torch.cumsum(torch.randint(1, 5000, (B, N)), dim=1)
    xin = torch.randn(B, N, D) * (torch.arange(N).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(-1)
    c = dict(max_seq=max_seq, out_len=out_len, D=D, blocks=blocks, H=H, dv=d, dqk=d, sd=proto.state_dict(), x=xin)
    args = dict(past_lengths=lengths.to(DEV), valid_mask=None, past_payloads={"timestamps": ts.to(DEV)})
    res = []
    for shared in (True, False):
        enc = _build(c, compute_dtype=torch.bfloat16)
        if not shared:
            monkeypatch.setattr(type(enc._hstu), "_layer_aux", lambda self, x: None)
        else:
            made = []
            orig = type(enc._hstu)._layer_aux
            monkeypatch.setattr(type(enc._hstu), "_layer_aux", lambda self, x: made.append(orig(self, x)) or made[-1])
        x = c["x"].to(DEV).requires_grad_(True)
        y, _ = enc(user_embeddings=x, **args)
        y.float().square().sum().backward()
        res.append((y.detach(), x.grad, {k: p.grad for k, p in enc.named_parameters() if p.grad is not None}))
        if shared:
            assert made and made[0] is not None and "dw_uvqk" in made[0][0] and made[0][0]["w_uvqk"].dtype == torch.bfloat16
        monkeypatch.undo()
    (y1, dx1, g1), (y2, dx2, g2) = res
    assert torch.equal(y1, y2) and torch.equal(dx1, dx2)
    assert g1.keys() == g2.keys() and len(g1) > 0
    for k in g1:
        _close(g1[k], g2[k], 1e-5, what=k)
    # the two helpers on their own
    ws = [torch.randn(257, 33, device=DEV), torch.randn(5, device=DEV), torch.randn(64, 1024, device=DEV)]
    for w, wc in zip(ws, GF.cast_many_bf16(ws)):
        assert wc.dtype == torch.bfloat16 and wc.data_ptr() % 16 == 0 and torch.equal(wc, w.bfloat16())
    zs = GF.zeros_many([(3, 5), (7,), (256, 1024)], torch.device(DEV))
    assert all(z.dtype == torch.float32 and not z.any() and z.data_ptr() % 256 == 0 for z in zs)
    assert [tuple(z.shape) for z in zs] == [(3, 5), (7,), (256, 1024)]


def test_training_mode_dropout_only_touches_o_input(golden):
    c = hstu_case(golden("hstu"), "mh")
    enc = _build(c).train()
    torch.manual_seed(0)
    y1, _ = enc(past_lengths=c["lengths"].to(DEV), user_embeddings=c["x"].to(DEV), valid_mask=None,
                past_payloads={"timestamps": c["ts"].to(DEV)})
    torch.manual_seed(0)
    y2, _ = enc(past_lengths=c["lengths"].to(DEV), user_embeddings=c["x"].to(DEV), valid_mask=None,
                past_payloads={"timestamps": c["ts"].to(DEV)})
    assert torch.equal(y1, y2) and not torch.allclose(y1.cpu(), c["y"], atol=1e-4)


# ---------------------------------------------------------------------------------------------
# tcgen05 path (bf16, head dim 64)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,H,lengths,with_ts", [
    (3, 211, 2, [211, 37, 129], True),              # HG = 2, two query tiles
    (6, 400, 4, [400, 129, 128, 1, 0, 257], True),  # HG = 4, tile-edge lengths, empty sequence
    (2, 300, 8, [300, 77], True),                   # two head groups
    (2, 260, 4, [260, 128], False),                 # no timestamps: no bias
    (1, 1024, 2, [1000], True),                     # eight key tiles: ring wrap-around
])
def test_attention_forward_bf16_tcgen05_vs_oracle(B, N, H, lengths, with_ts):
    d = 64
    c = _rand_case(B * 1000 + N, B, N, H, d, d, lengths, with_ts=with_ts)
    for nme in ("q", "k", "v"):
        c[nme] = c[nme].to(torch.bfloat16).float() * 0.5   # values exactly representable in bf16
    # q/k/v as strided views of one wide matrix, like torch.split of the uvqk output (hstu.py:308)
    wide = torch.cat([c["v"], c["q"], c["k"]], dim=1).to(DEV).to(torch.bfloat16)
    W = H * d
    vg, qg, kg = wide[:, :W], wide[:, W:2 * W], wide[:, 2 * W:]
    ts = c["ts"].to(DEV) if with_ts else None
    out = GF.hstu_attention(qg, kg, vg, c["off"].to(DEV), ts,
                            c["ts_w"].to(DEV) if with_ts else None,
                            c["pos_w"].to(DEV) if with_ts else None,
                            _thr() if with_ts else None, N, H, d, d)
    assert out.dtype == torch.bfloat16
    ref = O.hstu_attention(c["q"].double(), c["k"].double(), c["v"].double(), c["off"], c["ts"],
                           c["ts_w"].double() if with_ts else None,
                           c["pos_w"].double() if with_ts else None, N, H, d, d)
    _close(out, ref, 1e-2, 5e-3, what="tcgen05 attn fwd")


def test_bf16_tcgen05_large_time_gaps_take_the_exact_slow_path():
    # millisecond-style timestamps: gaps beyond 2^32 use the 64-bit binary search
    B, N, H, d = 2, 200, 2, 64
    c = _rand_case(77, B, N, H, d, d, [200, 150])
    c["ts"] = c["ts"] * 100_000 * (c["ts"] > 0)
    for nme in ("q", "k", "v"):
        c[nme] = c[nme].to(torch.bfloat16).float() * 0.5
    out, _ = _run_kernel(c, N, H, d, d, dtype=torch.bfloat16)
    ref = O.hstu_attention(c["q"].double(), c["k"].double(), c["v"].double(), c["off"], c["ts"],
                           c["ts_w"].double(), c["pos_w"].double(), N, H, d, d)
    _close(out, ref, 1e-2, 5e-3, what="tcgen05 attn fwd, 64-bit gaps")


@pytest.mark.parametrize("B,N,H,lengths,with_ts,scale_ts", [
    (3, 211, 2, [211, 37, 129], True, 1),
    (6, 400, 4, [400, 129, 128, 1, 0, 257], True, 1),
    (2, 300, 8, [300, 77], True, 1),
    (2, 260, 4, [260, 128], False, 1),
    (1, 700, 2, [650], True, 1),
    (2, 200, 2, [200, 150], True, 100_000),      # gaps beyond 2^32: exact 64-bit bucketing path
])
def test_attention_backward_bf16_tcgen05_vs_oracle(B, N, H, lengths, with_ts, scale_ts):
    d = 64
    c = _rand_case(B * 1000 + N + 1, B, N, H, d, d, lengths, with_ts=with_ts)
    if with_ts and scale_ts != 1:
        c["ts"] = c["ts"] * scale_ts
    for nme in ("q", "k", "v"):
        c[nme] = c[nme].to(torch.bfloat16).float() * 0.5
    gen = torch.Generator().manual_seed(2)
    w = torch.randn(c["T"], H * d, generator=gen).to(torch.bfloat16).float()
    out, leaves = _run_kernel(c, N, H, d, d, dtype=torch.bfloat16, grad=True)
    out.backward(w.to(DEV).to(torch.bfloat16))
    ref_leaves = [c[n].clone().double().requires_grad_(True) for n in ("q", "k", "v", "ts_w", "pos_w")]
    ref = O.hstu_attention(ref_leaves[0], ref_leaves[1], ref_leaves[2], c["off"], c["ts"],
                           ref_leaves[3] if with_ts else None, ref_leaves[4] if with_ts else None,
                           N, H, d, d)
    (ref * w.double()).sum().backward()
    names = ("dq", "dk", "dv") + (("d_ts_w", "d_pos_w") if with_ts else ())
    for name, got, r in zip(names, leaves, ref_leaves):
        _close(got.grad, r.grad, 2e-2, 2e-2, what=f"tcgen05 {name}")


@pytest.mark.parametrize("B,N,H,lengths,scale_ts", [
    (3, 400, 4, [400, 129, 257], 1),
    (2, 300, 2, [300, 77], 100_000),           # buckets through the 64-bit path
])
def test_bucket_cache_gives_the_same_forward_and_backward(B, N, H, lengths, scale_ts):
    d = 64
    c = _rand_case(5 + N, B, N, H, d, d, lengths)
    c["ts"] = c["ts"] * scale_ts
    res = []
    for use_cache in (False, True):
        q, k, v = (c[n].to(DEV).to(torch.bfloat16).requires_grad_(True) for n in ("q", "k", "v"))
        ts_w = c["ts_w"].to(DEV).requires_grad_(True)
        pos_w = c["pos_w"].to(DEV).requires_grad_(True)
        off, ts = c["off"].to(DEV), c["ts"].to(DEV)
        cache = GF.hstu_bucket_cache(off, ts, _thr(), N) if use_cache else None
        out = GF.hstu_attention(q, k, v, off, ts, ts_w, pos_w, _thr(), N, H, d, d, bucket_cache=cache)
        out.backward(torch.ones_like(out))
        res.append((out.detach(), q.grad, k.grad, v.grad, ts_w.grad, pos_w.grad))
    # two different forward kernels (buckets derived in the kernel, f16 SiLU chain / bucket cache, bf16
    # SiLU chain): same buckets, outputs equal within the bf16 tolerance
    _close(res[1][0], res[0][0], 1e-2, 5e-3, what="cached fwd")
    for a, b_ in zip(res[0][1:4], res[1][1:4]):
        _close(b_, a, 1e-3, what="cached bwd")                   # dq: fp32 atomics reorder
    _close(res[1][4], res[0][4], 2e-3, what="cached d_ts_w")
    _close(res[1][5], res[0][5], 2e-3, what="cached d_pos_w")


def test_bucket_cache_contents_match_the_reference_bucketization():
    B, N = 2, 300
    c = _rand_case(9, B, N, 2, 64, 64, [300, 140])
    off, ts = c["off"].to(DEV), c["ts"].to(DEV)
    cache = GF.hstu_bucket_cache(off, ts, _thr(), N).cpu().view(B, -1, 2, 8, 128, 16)
    ext = torch.cat([c["ts"], c["ts"][:, N - 1:N]], dim=1)
    ref = O.bucketize_ts(ext[:, 1:].unsqueeze(2) - ext[:, :-1].unsqueeze(1))     # (B, N, N)
    for b, n in enumerate([300, 140]):
        for iq in range((n + 127) // 128):
            for jk in range(iq + 1):
                slot = iq * (iq + 1) // 2 + jk
                # Q orientation: [chunk][row][16] -> (row, col)
                tile_q = cache[b, slot, 0].permute(1, 0, 2).reshape(128, 128)
                tile_k = cache[b, slot, 1].permute(1, 0, 2).reshape(128, 128).t()
                rows = min(128, N - iq * 128)
                cols = min(128, N - jk * 128)
                exp = ref[b, iq * 128: iq * 128 + rows, jk * 128: jk * 128 + cols].to(torch.uint8)
                assert torch.equal(tile_q[:rows, :cols], exp)
                assert torch.equal(tile_k[:rows, :cols], exp)


def test_tcgen05_stress_deterministic_and_matches_cuda_core_path(monkeypatch):
    """Race hunting without a sanitizer: random ragged shapes, every launch repeated; the forward
    has no atomics so it must be bit-reproducible, and both directions must agree with the
    CUDA-core kernels run on the same bf16 inputs."""
    gen = torch.Generator().manual_seed(123)
    d = 64
    for trial in range(12):
        B = int(torch.randint(1, 5, (1,), generator=gen))
        H = [2, 4, 8][trial % 3]
        N = int(torch.randint(130, 700, (1,), generator=gen))
        lengths = torch.randint(0, N + 1, (B,), generator=gen).tolist()
        lengths[0] = N
        c = _rand_case(1000 + trial, B, N, H, d, d, lengths)
        for nme in ("q", "k", "v"):
            c[nme] = c[nme].to(torch.bfloat16).float() * 0.5
        w = torch.randn(c["T"], H * d, generator=gen).to(DEV).to(torch.bfloat16)
        cache = GF.hstu_bucket_cache(c["off"].to(DEV), c["ts"].to(DEV), _thr(), N) if trial % 2 else None

        def run():
            q, k, v = (c[n].to(DEV).to(torch.bfloat16).requires_grad_(True) for n in ("q", "k", "v"))
            ts_w = c["ts_w"].to(DEV).requires_grad_(True)
            pos_w = c["pos_w"].to(DEV).requires_grad_(True)
            out = GF.hstu_attention(q, k, v, c["off"].to(DEV), c["ts"].to(DEV), ts_w, pos_w, _thr(),
                                    N, H, d, d, bucket_cache=cache)
            out.backward(w)
            return out.detach(), q.grad, k.grad, v.grad, ts_w.grad, pos_w.grad

        first = run()
        for _ in range(2):
            again = run()
            assert torch.equal(first[0], again[0]), f"trial {trial}: forward not reproducible"
            assert torch.equal(first[2], again[2]) and torch.equal(first[3], again[3]), \
                f"trial {trial}: dK/dV not reproducible"          # no atomics on dK, dV either
        monkeypatch.setenv("GRB_FORCE_CUDA_CORE", "1")
        ref = run()
        monkeypatch.delenv("GRB_FORCE_CUDA_CORE")
        _close(first[0], ref[0], 2e-2, 1e-2, what=f"trial {trial} out")
        for name, a, b_ in zip(("dq", "dk", "dv", "d_ts_w", "d_pos_w"), first[1:], ref[1:]):
            _close(a, b_, 3e-2, 2e-2, what=f"trial {trial} {name}")


# ---------------------------------------------------------------------------------------------
# incremental path: delta_x_offsets + cache (hstu.py:151-177, :293-298, :321-322, :397-401, :415-418)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,H,d,with_ts", [
    (torch.float32, 2, 8, True), (torch.float32, 1, 50, True), (torch.float32, 2, 64, False),
    (torch.bfloat16, 2, 64, True), (torch.bfloat16, 4, 64, False), (torch.bfloat16, 1, 24, True),
])
def test_attention_decode_equals_row_of_full_attention(dtype, H, d, with_ts):
    """grb_hstu_attn_decode against the oracle's padded attention (all rows computed, one kept),
    at the last position and at positions in the middle of the sequences."""
    B, N = 5, 211
    lengths = [211, 1, 130, 37, 64]
    c = _rand_case(7, B, N, H, d, d, lengths, with_ts=with_ts)
    q, k, v = (c[n].to(dtype).float() for n in ("q", "k", "v"))      # the values the kernel sees
    ref = O.hstu_attention(q, k, v, c["off"], c["ts"], c["ts_w"], c["pos_w"], N, H, d, d)
    for positions in (c["lengths"] - 1, torch.tensor([100, 0, 0, 36, 31]), (c["lengths"] - 1).to(torch.int32)):
        rows = c["off"][:-1] + positions.long()
        kc = O.jagged_to_padded_dense(k, c["off"], N, 0.0)
        got = GF.hstu_attention_decode(
            q[rows].to(DEV).to(dtype), kc.to(DEV).to(dtype), v.to(DEV).to(dtype), c["off"].to(DEV),
            positions.to(DEV), c["ts"].to(DEV) if with_ts else None,
            c["ts_w"].to(DEV) if with_ts else None, c["pos_w"].to(DEV) if with_ts else None,
            _thr() if with_ts else None, N, H, d, d)
        assert got.shape == (B, H * d) and got.dtype == dtype
        if dtype == torch.float32:
            _close(got, ref[rows], 1e-5, what="decode fp32")
        else:   # inputs identical, fp32 accumulation: only the bf16 rounding of the output differs
            _close(got, ref[rows], 1e-2, 5e-3, what="decode bf16")


def test_attention_decode_rejects_autograd_and_cpu():
    q = torch.zeros(1, 8, device=DEV, requires_grad=True)
    kc = torch.zeros(1, 4, 8, device=DEV)
    v = torch.zeros(2, 8, device=DEV)
    off = torch.tensor([0, 2], device=DEV)
    pos = torch.tensor([1], device=DEV)
    with pytest.raises(NotImplementedError):
        GF.hstu_attention_decode(q, kc, v, off, pos, None, None, None, None, 4, 1, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GF.hstu_attention_decode(q.detach().cpu(), kc, v, off, pos, None, None, None, None, 4, 1, 8, 8)


def _inc_build(c, compute_dtype=None):
    return _build(dict(c, sd=c["sd"]), compute_dtype=compute_dtype)


@pytest.mark.parametrize("name", ["mh", "h64"])
def test_hstu_incremental_fp32_vs_reference_golden(golden, name):
    c = hstu_incremental_case(golden("hstu_incremental"), name)
    enc = _inc_build(c)
    kw = dict(past_lengths=c["lengths"].to(DEV), valid_mask=None,
              past_payloads={"timestamps": c["ts"].to(DEV)})
    with torch.no_grad():
        y0, cache = enc(user_embeddings=c["x"].to(DEV), return_cache_states=True, **kw)
        _close(y0, c["y0"], 1e-5, what=f"{name} y0")
        assert len(cache) == c["blocks"]
        for got, ref in zip(cache, c["cache0"]):
            for a, b, nm in zip(got, ref, ("v", "padded_q", "padded_k", "out")):
                assert a.shape == b.shape
                _close(a, b, 1e-5, what=f"{name} cache {nm}")
        # (1) from the reference's own cache states, (2) from ours; delta indices int64 and int32
        for states, cast in (([tuple(t.to(DEV).clone() for t in st) for st in c["cache0"]], torch.int64),
                             (cache, torch.int32)):
            delta = tuple(t.to(DEV).to(cast) for t in c["delta"])
            y, new_states = enc(user_embeddings=c["x2"].to(DEV), delta_x_offsets=delta, cache=states,
                                return_cache_states=True, **kw)
            _close(y, c["y_inc"], 1e-5, what=f"{name} y_inc")
            assert len(new_states) == c["blocks"]
            # in place, like the reference (index_copy_): the caller's tensors now hold the new rows
            assert new_states[0][0].data_ptr() == states[0][0].data_ptr()
            if c["cache1"] is not None:
                for got, ref in zip(new_states, c["cache1"]):
                    for a, b, nm in zip(got, ref, ("v", "padded_q", "padded_k", "out")):
                        _close(a, b, 1e-5, what=f"{name} updated cache {nm}")


def test_hstu_incremental_bf16_matches_full_pass(golden):
    """bf16 / tcgen05 configuration: recomputing the last token through the caches equals a full
    pass over the modified sequence (and the reference's fp32 result within the bf16 tolerance)."""
    c = hstu_incremental_case(golden("hstu_incremental"), "h64")
    enc = _inc_build(c, compute_dtype=torch.bfloat16)
    kw = dict(past_lengths=c["lengths"].to(DEV), valid_mask=None,
              past_payloads={"timestamps": c["ts"].to(DEV)})
    with torch.no_grad():
        _, cache = enc(user_embeddings=c["x"].to(DEV), return_cache_states=True, **kw)
        assert cache[0][1].dtype == torch.bfloat16
        delta = tuple(t.to(DEV) for t in c["delta"])
        y_inc, _ = enc(user_embeddings=c["x2"].to(DEV), delta_x_offsets=delta, cache=cache, **kw)
        y_full, _ = enc(user_embeddings=c["x2"].to(DEV), **kw)
    _close(y_inc, y_full, 1e-2, 5e-3, what="bf16 incremental vs full")
    _close(y_inc, c["y_inc"], 2e-2, 1e-2, what="bf16 incremental vs reference")


def test_hstu_incremental_needs_cache(golden):
    c = hstu_incremental_case(golden("hstu_incremental"), "mh")
    enc = _inc_build(c)
    with pytest.raises(ValueError, match="cache"), torch.no_grad():
        enc(past_lengths=c["lengths"].to(DEV), user_embeddings=c["x2"].to(DEV), valid_mask=None,
            past_payloads={"timestamps": c["ts"].to(DEV)},
            delta_x_offsets=tuple(t.to(DEV) for t in c["delta"]))


# ---------------------------------------------------------------------------------------------
# normalization="softmax_rel_bias" (hstu.py:337-384): composite of the jagged kernels, cuBLAS, ATen
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["mh", "h64"])
def test_hstu_softmax_rel_bias_vs_reference_golden(golden, name):
    c = hstu_case(golden("hstu_softmax"), name)
    enc = _build(c, normalization="softmax_rel_bias")
    x = c["x"].to(DEV).requires_grad_(True)
    y, _ = enc(past_lengths=c["lengths"].to(DEV), user_embeddings=x, valid_mask=None,
               past_payloads={"timestamps": c["ts"].to(DEV)})
    _close(y, c["y"], 1e-5, what=f"softmax {name} y")
    (y * c["w"].to(DEV)).sum().backward()
    _close(x.grad, c["dx"], 2e-4, what=f"softmax {name} dx")
    for k, p in enc.named_parameters():
        _close(p.grad, c["grads"][k], 5e-4, what=f"softmax {name} grad {k}")
    with pytest.raises(NotImplementedError, match="incremental"), torch.no_grad():
        _, cache = enc(past_lengths=c["lengths"].to(DEV), user_embeddings=x.detach(), valid_mask=None,
                       past_payloads={"timestamps": c["ts"].to(DEV)}, return_cache_states=True)
        off = O.complete_cumsum(c["lengths"])
        enc(past_lengths=c["lengths"].to(DEV), user_embeddings=x.detach(), valid_mask=None,
            past_payloads={"timestamps": c["ts"].to(DEV)}, cache=cache,
            delta_x_offsets=((off[1:] - 1).to(DEV), (c["lengths"] - 1).to(DEV)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_silu_split_matches_torch(dtype):
    """grb_silu_fwd / grb_silu_split_bwd against F.silu + torch.split (hstu.py:304-320); one of the
    pieces gets no gradient, one arrives as a strided view."""
    gen = torch.Generator().manual_seed(3)
    rows, sizes = 1037, [64, 64, 32, 96]
    x = (torch.randn(rows, sum(sizes), generator=gen) * 3).to(DEV).to(dtype)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    pa = GF.silu_split(xa, sizes)
    pb = torch.split(torch.nn.functional.silu(xb), sizes, dim=1)
    for a, b in zip(pa, pb):
        assert a.shape == b.shape
        torch.testing.assert_close(a, b, rtol=1e-6 if dtype == torch.float32 else 1e-2, atol=1e-6)
    big = torch.randn(rows, 2 * sizes[3], generator=gen).to(DEV).to(dtype)
    w0 = torch.randn(rows, sizes[0], generator=gen).to(DEV).to(dtype)
    w2 = torch.randn(rows, sizes[2], generator=gen).to(DEV).to(dtype)
    for p, xx in ((pa, xa), (pb, xb)):      # piece 1 unused, piece 3 multiplied by a strided view
        ((p[0] * w0).sum() + (p[2] * w2).sum() + (p[3] * big[:, ::2]).sum()).backward()
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(xa.grad, xb.grad, **tol)
    assert xa.grad[:, sizes[0]:sizes[0] + sizes[1]].abs().max().item() == 0.0


# ---------------------------------------------------------------------------------------------
# layer options no shipped config switches on: concat_ua (hstu.py:398-402),
# linear_activation="none" (:304-307), no relative attention bias
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kw,with_ts", [
    ("ua", dict(concat_ua=True), True), ("ua64", dict(concat_ua=True), True),
    ("noact", dict(linear_activation="none"), True),
    ("norab", dict(enable_relative_attention_bias=False), False),
])
def test_hstu_layer_options_vs_reference_golden(golden, name, kw, with_ts):
    c = hstu_case(golden("hstu_options"), name)
    enc = _build(c, **kw)
    x = c["x"].to(DEV).requires_grad_(True)
    y, _ = enc(past_lengths=c["lengths"].to(DEV), user_embeddings=x, valid_mask=None,
               past_payloads={"timestamps": c["ts"].to(DEV)} if with_ts else {})
    _close(y, c["y"], 1e-5, what=f"{name} y")
    (y * c["w"].to(DEV)).sum().backward()
    _close(x.grad, c["dx"], 2e-4, what=f"{name} dx")
    for k, p in enc.named_parameters():
        _close(p.grad, c["grads"][k], 5e-4, what=f"{name} grad {k}")
    if name == "ua64":      # bf16 / tcgen05 configuration of the same case
        enc16 = _build(c, compute_dtype=torch.bfloat16, **kw)
        y16, _ = enc16(past_lengths=c["lengths"].to(DEV), user_embeddings=c["x"].to(DEV), valid_mask=None,
                       past_payloads={"timestamps": c["ts"].to(DEV)})
        _close(y16, c["y"], 2e-2, 1e-2, what="ua64 bf16 y")


@pytest.mark.parametrize("H,d,bias_kind", [(1, 256, "bucketed"), (1, 160, "bucketed"), (2, 64, "positional")])
def test_wide_heads_and_positional_bias_train_through_the_composite(H, d, bias_kind):
    """The reference's default injection (attention_dim = linear_dim = item_embedding_dim, i.e. one head of
    256, generative_recommenders.py:157-160) and RelativePositionalBias (hstu.py:50-68) are outside the
    fused kernels' shapes: the layer takes the reference's padded formulation on GPU ops instead.  Forward
    and all gradients against the CPU oracle of the same layer (fp32, 1e-4)."""
    torch.manual_seed(3)
    B, N, D = 4, 50, 64
    lengths = torch.tensor([50, 7, 33, 1])
    off = torch.zeros(B + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lengths, 0)
    T = int(off[-1])
    bias = (hstu.RelativeBucketedTimeAndPositionBasedBias(N, 128, hstu._default_bucketization)
            if bias_kind == "bucketed" else hstu.RelativePositionalBias(N))
    layer = hstu.SequentialTransductionUnitJagged(
        embedding_dim=D, linear_hidden_dim=d, attention_dim=d, dropout_ratio=0.0, attn_dropout_ratio=0.0,
        num_heads=H, linear_activation="silu", relative_attention_bias_module=bias, max_length=N)
    x = torch.randn(T, D)
    ts = 978_300_000 + torch.cumsum(torch.randint(1, 5000, (B, N)), dim=1)
    ts = ts * (torch.arange(N).unsqueeze(0) <= lengths.unsqueeze(1))
    mask = torch.ones(N, N).tril_()
    # CPU oracle: the same padded formulation, float64
    sd = {k: v.detach().double().requires_grad_(True) for k, v in layer.state_dict().items()}
    xr = x.double().requires_grad_(True)
    xn = torch.nn.functional.layer_norm(xr, (D,), eps=1e-6)
    u, v, q, k = torch.split(torch.nn.functional.silu(xn @ sd["_uvqk"]), [d * H, d * H, d * H, d * H], dim=1)
    pad = lambda t: O.jagged_to_padded_dense(t, off, N)
    qk = torch.einsum("bnhd,bmhd->bhnm", pad(q).view(B, N, H, d), pad(k).view(B, N, H, d))
    if bias_kind == "bucketed":
        rb = O.rel_bias(ts, sd["_rel_attn_bias._ts_w"], sd["_rel_attn_bias._pos_w"], N)
    else:
        ar = torch.arange(N)
        rb = sd["_rel_attn_bias._w"][(N - 1) + ar.view(1, N) - ar.view(N, 1)].unsqueeze(0)
    qk = torch.nn.functional.silu(qk + rb.unsqueeze(1)) / N * mask.double()
    a = O.dense_to_jagged(torch.einsum("bhnm,bmhd->bnhd", qk, pad(v).view(B, N, H, d)).reshape(B, N, H * d), off)
    ref = (u * torch.nn.functional.layer_norm(a, (H * d,), eps=1e-6)) @ sd["_o.weight"].t() + sd["_o.bias"] + xr
    w = torch.randn(T, D)
    (ref * w.double()).sum().backward()
    layer = layer.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y, _ = layer(xg, off.to(DEV), ts.to(DEV), mask.to(DEV))
    (y * w.to(DEV)).sum().backward()
    _close(y, ref, 1e-4, what="composite fwd")
    _close(xg.grad, xr.grad, 1e-4, what="composite dx")
    for name, p in layer.named_parameters():
        _close(p.grad, sd[name].grad, 2e-4, what=f"composite d{name}")


def test_non_causal_mask_is_refused_not_ignored():
    enc = hstu.HSTU(max_sequence_len=20, max_output_len=4, embedding_dim=64, item_embedding_dim=64, num_blocks=1,
                    num_heads=1, linear_dim=64, attention_dim=64, normalization="rel_bias", linear_config="uvqk",
                    linear_activation="silu", linear_dropout_rate=0.0, attn_dropout_rate=0.0).to(DEV).eval()
    N = 24
    lengths = torch.tensor([10, 24], device=DEV)
    x = torch.randn(2, N, 64, device=DEV)
    ts = torch.zeros(2, N, dtype=torch.int64, device=DEV)
    enc(past_lengths=lengths, user_embeddings=x, valid_mask=None, past_payloads={"timestamps": ts})   # causal: fine
    full = torch.ones(N, N, device=DEV)
    xj = torch.randn(34, 64, device=DEV)
    off = torch.tensor([0, 10, 34], device=DEV)
    with pytest.raises(NotImplementedError, match="causal"):
        enc._hstu.jagged_forward(xj, off, ts, full)
    enc._hstu.jagged_forward(xj, off, ts, 1.0 - enc._attn_mask.float())      # the reference's spelling: fine
