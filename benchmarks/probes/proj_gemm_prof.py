"""The six projection GEMMs of one HSTU layer at the C2 shape (T = 14 082 rows, D = 256, 4 x 64
heads), a few times each, for ncu / timing: python benchmarks/probes/proj_gemm_prof.py [iters]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from mygenerativerecommenders_b200 import _lib  # noqa: E402
from mygenerativerecommenders_b200 import functional as GF  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda"
T, D, N = 14082, 256, 1024
bf = lambda *s: (torch.randn(*s, device=dev) * 0.3).to(torch.bfloat16)
xn, w_uvqk, dx = bf(T, D), bf(D, N), bf(T, N)
o_in, w_o, g, res = bf(T, D), bf(D, D), bf(T, D), bf(T, D)
bias = torch.randn(D, device=dev)
pre, act = torch.empty(T, N, dtype=torch.bfloat16, device=dev), torch.empty(T, N, dtype=torch.bfloat16, device=dev)
out = torch.empty(T, D, dtype=torch.bfloat16, device=dev)
dw1 = torch.zeros(D, N, device=dev)
dw2 = torch.zeros(D, D, device=dev)
calls = {
    "uvqk_fwd": lambda: GF._proj_gemm(xn, w_uvqk, T, N, D, False, True, _lib.GEMM_EPI_SILU2, pre, act),
    "o_fwd": lambda: GF._proj_gemm(o_in, w_o, T, D, D, False, False, _lib.GEMM_EPI_BIAS_RES, out, bias=bias, res=res),
    "uvqk_dgrad": lambda: GF._proj_gemm(dx, w_uvqk, T, D, N, False, False, _lib.GEMM_EPI_PLAIN, out),
    "o_dgrad": lambda: GF._proj_gemm(g, w_o, T, D, D, False, True, _lib.GEMM_EPI_PLAIN, out),
    "uvqk_wgrad": lambda: GF._proj_gemm(xn, dx, D, N, T, True, True, _lib.GEMM_EPI_F32_ADD, dw1),
    "o_wgrad": lambda: GF._proj_gemm(g, o_in, D, D, T, True, True, _lib.GEMM_EPI_F32_ADD, dw2),
}
for name, fn in calls.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(4_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us")
# cuBLAS for comparison (library)
for name, fn in {"cublas uvqk_fwd": lambda: torch.mm(xn, w_uvqk), "cublas uvqk_wgrad": lambda: torch.mm(xn.t(), dx, out_dtype=torch.float32),
                 "cublas uvqk_dgrad": lambda: torch.mm(dx, w_uvqk.t()), "cublas o_fwd": lambda: torch.addmm(bias.to(torch.bfloat16), o_in, w_o.t())}.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(4_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us")
