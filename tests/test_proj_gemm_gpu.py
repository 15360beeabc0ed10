"""f1: the tcgen05 projection GEMM with fused epilogues (csrc/proj_gemm.cu) against a plain PyTorch
fp32 reference of the same op on the same bf16 inputs (hstu.py:302-320, :404-413 and their backward).
Tolerance: the outputs are bf16 roundings of fp32-accumulated sums: |d| <= 1e-2 * max|ref| elementwise,
rel-L2 <= 5e-3 (DESIGN.md §2, bf16 path); the fp32 weight-gradient epilogue: rel-L2 <= 1e-4."""
import pytest
import torch

from mygenerativerecommenders_b200 import _lib
from mygenerativerecommenders_b200 import functional as GF

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, ref, tol_inf, tol_l2, what):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    scale = max(ref.abs().max().item(), 1e-12)
    err = (got - ref).abs().max().item()
    assert err <= tol_inf * scale, f"{what}: max|d|={err:.3e} > {tol_inf:.0e}*{scale:.3e}"
    rel = ((got - ref).norm() / max(ref.norm().item(), 1e-12)).item()
    assert rel <= tol_l2, f"{what}: rel-l2 {rel:.3e} > {tol_l2:.0e}"


def _bf(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


@pytest.mark.parametrize("M,N,K", [(14082, 1024, 256), (130, 256, 1024), (1, 256, 64), (4096, 512, 192)])
@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
def test_plain_epilogue_every_operand_layout(M, N, K, a_mn, b_mn):
    if a_mn and M % 8:
        pytest.skip("a transposed A needs a 16-byte row stride: M % 8 == 0")
    A = _bf(M, K, seed=1, scale=0.5)
    B = _bf(K, N, seed=2, scale=0.5)
    ref = A.float() @ B.float()
    a_st = A.t().contiguous() if a_mn else A              # stored (K, M) or (M, K)
    b_st = B if b_mn else B.t().contiguous()              # stored (K, N) or (N, K)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    GF._proj_gemm(a_st, b_st, M, N, K, a_mn, b_mn, _lib.GEMM_EPI_PLAIN, out)
    _close(out, ref, 1e-2, 5e-3, f"plain a_mn={a_mn} b_mn={b_mn}")


def test_silu2_epilogue_writes_pre_activation_and_its_silu():
    M, N, K = 14082, 1024, 256
    A, B = _bf(M, K, seed=3), _bf(K, N, seed=4, scale=0.1)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    act = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    GF._proj_gemm(A, B, M, N, K, False, True, _lib.GEMM_EPI_SILU2, pre, act)
    ref = A.float() @ B.float()
    _close(pre, ref, 1e-2, 5e-3, "pre-activation")
    # the activation is SiLU of the STORED (bf16) pre-activation
    _close(act, torch.nn.functional.silu(pre.float()), 1e-2, 5e-3, "silu(pre)")


def test_bias_residual_epilogue():
    M, N, K = 5000, 256, 256
    A, W = _bf(M, K, seed=5), _bf(N, K, seed=6, scale=0.1)       # W stored (N, K): nn.Linear layout
    res = _bf(M, N, seed=7)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(8)).to(DEV)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    GF._proj_gemm(A, W, M, N, K, False, False, _lib.GEMM_EPI_BIAS_RES, out, bias=bias, res=res)
    ref = A.float() @ W.float().t() + bias + res.float()
    _close(out, ref, 1e-2, 5e-3, "bias + residual")


@pytest.mark.parametrize("T", [14082, 777, 64, 33])
def test_weight_gradient_split_k_fp32(T):
    D, N = 256, 1024
    X, G = _bf(T, D, seed=9), _bf(T, N, seed=10, scale=0.2)
    dw = torch.zeros(D, N, dtype=torch.float32, device=DEV)
    GF._proj_gemm(X, G, D, N, T, True, True, _lib.GEMM_EPI_F32_ADD, dw)
    ref = X.float().t() @ G.float()
    _close(dw, ref, 1e-4, 1e-4, "wgrad")
    GF._proj_gemm(X, G, D, N, T, True, True, _lib.GEMM_EPI_F32_ADD, dw)     # accumulates
    _close(dw, 2 * ref, 1e-4, 1e-4, "wgrad accumulate")


def test_colsum_bias_gradient():
    T, W = 14082, 256
    G = _bf(T, W, seed=11)
    out = torch.zeros(W, dtype=torch.float32, device=DEV)
    _lib.check(_lib.lib().grb_colsum_bf16(G.data_ptr(), G.stride(0), T, W, out.data_ptr(),
                                          _lib.stream_ptr(torch.device(DEV))))
    _close(out, G.float().sum(0), 1e-4, 1e-4, "colsum")


def _layer_pair(T=3000, D=256, Ntot=1024):
    g = torch.Generator().manual_seed(12)
    xn = (torch.randn(T, D, generator=g)).to(torch.bfloat16).to(DEV)
    w = (torch.randn(D, Ntot, generator=g) * 0.05).to(DEV)
    return xn, w


def test_uvqk_projection_autograd_vs_torch():
    xn, w = _layer_pair()
    sizes = [256, 256, 256, 256]
    gs = [_bf(xn.shape[0], 256, seed=20 + i) for i in range(4)]
    x1, w1 = xn.clone().requires_grad_(True), w.clone().requires_grad_(True)
    parts = GF.uvqk_projection(x1, w1, sizes)
    torch.autograd.backward(parts, gs)
    x2, w2 = xn.float().clone().requires_grad_(True), w.clone().requires_grad_(True)
    # reference on the bf16-rounded weight (what the kernel multiplies by), fp32 math
    ref = torch.split(torch.nn.functional.silu(x2 @ w2.to(torch.bfloat16).float()), sizes, dim=1)
    torch.autograd.backward(ref, [g_.float() for g_ in gs])
    for a, b in zip(parts, ref):
        _close(a, b, 1e-2, 5e-3, "uvqk fwd")
    _close(x1.grad, x2.grad, 2e-2, 1e-2, "uvqk dgrad")
    _close(w1.grad, w2.grad, 2e-2, 1e-2, "uvqk wgrad")
    assert w1.grad.dtype == torch.float32


def test_output_projection_autograd_vs_torch():
    T, D = 3000, 256
    g = torch.Generator().manual_seed(13)
    o_in = torch.randn(T, D, generator=g).to(torch.bfloat16).to(DEV)
    res = torch.randn(T, D, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(D, D, generator=g) * 0.05).to(DEV)
    b = torch.randn(D, generator=g).to(DEV)
    go = _bf(T, D, seed=30)
    leaves = [t.clone().requires_grad_(True) for t in (o_in, w, b, res)]
    out = GF.output_projection(*leaves)
    out.backward(go)
    r = [o_in.float().clone().requires_grad_(True), w.clone().requires_grad_(True),
         b.clone().requires_grad_(True), res.float().clone().requires_grad_(True)]
    ref = torch.nn.functional.linear(r[0], r[1].to(torch.bfloat16).float(), r[2]) + r[3]
    ref.backward(go.float())
    _close(out, ref, 1e-2, 5e-3, "o fwd")
    for name, a, b_ in zip(("d_in", "dW", "db", "d_res"), leaves, r):
        _close(a.grad, b_.grad, 2e-2, 1e-2, f"o {name}")
