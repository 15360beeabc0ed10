#!/usr/bin/env python
"""Per-kernel micro-benchmarks with roofline fractions (CUDA events, L2 flushed between timed
iterations).  One JSON object per line.  Usage: python benchmarks/kbench.py [names...]"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mygenerativerecommenders_b200 import functional as GF  # noqa: E402
from mygenerativerecommenders_b200 import hstu, ops  # noqa: E402

DEV = torch.device("cuda")
PK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def timeit(fn, iters=10, warmup=3, flush=True):
    import os
    if os.environ.get("KBENCH_FAST"):
        iters, warmup = 1, 1
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def report(name, ms, flop=None, byts=None, **kw):
    out = {"kernel": name, "ms": round(ms, 4)}
    if flop is not None:
        out["tflops"] = round(flop / ms / 1e9, 2)
        out["tensor_frac"] = round(flop / ms / 1e9 / PK["bf16_tflops"], 4)
    if byts is not None:
        out["gbs"] = round(byts / ms / 1e6, 1)
        out["hbm_frac"] = round(byts / ms / 1e6 / PK["hbm_gbs"], 4)
    out.update(kw)
    print(json.dumps(out), flush=True)


def attn_case(B, N, H, lengths, dtype=torch.bfloat16, d=64):
    lengths = torch.as_tensor(lengths, dtype=torch.int64)
    off = torch.zeros(B + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lengths, 0)
    T = int(off[-1])
    g = torch.Generator(device=DEV).manual_seed(0)
    mk = lambda w: (torch.randn(T, w, device=DEV, generator=g) * 0.5).to(dtype)
    q, k, v = mk(H * d), mk(H * d), mk(H * d)
    ts = 978_300_000 + torch.cumsum(torch.randint(1, 5000, (B, N)), 1)
    ts = (ts * (torch.arange(N).unsqueeze(0) <= lengths.unsqueeze(1))).to(DEV)
    ts_w = (torch.randn(129, device=DEV) * 0.02)
    pos_w = (torch.randn(2 * N - 1, device=DEV) * 0.02)
    thr = hstu.tabulate_bucket_thresholds(hstu._default_bucketization, 128).to(DEV)
    pairs = int((lengths * (lengths + 1) // 2).sum())
    return dict(q=q, k=k, v=v, off=off.to(DEV), ts=ts, ts_w=ts_w, pos_w=pos_w, thr=thr, N=N, H=H,
                d=d, pairs=pairs, T=T)


def bench_attn(tag, B, N, H, lengths, bwd=True):
    c = attn_case(B, N, H, lengths)
    H, d = c["H"], c["d"]
    mkc = lambda: GF.hstu_bucket_cache(c["off"], c["ts"], c["thr"], c["N"])
    cache = mkc()
    report(f"hstu_bucket_tiles[{tag}] (once per batch, shared by all layers)", timeit(mkc, flush=False),
           byts=cache.numel())
    run = lambda: GF.hstu_attention(c["q"], c["k"], c["v"], c["off"], c["ts"], c["ts_w"],
                                    c["pos_w"], c["thr"], c["N"], H, d, d, bucket_cache=cache)
    ms = timeit(run, flush=False)
    report(f"hstu_attn_fwd[{tag}]", ms, flop=c["pairs"] * 2 * H * 2 * d, T=c["T"], pairs=c["pairs"])
    if bwd:
        q, k, v = (c[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
        ts_w, pos_w = c["ts_w"].clone().requires_grad_(True), c["pos_w"].clone().requires_grad_(True)
        out = GF.hstu_attention(q, k, v, c["off"], c["ts"], ts_w, pos_w, c["thr"], c["N"], H, d, d,
                                bucket_cache=cache)
        go = torch.randn_like(out)
        runb = lambda: torch.autograd.grad(out, (q, k, v, ts_w, pos_w), go, retain_graph=True)
        msb = timeit(runb, iters=5, warmup=2, flush=False)
        report(f"hstu_attn_bwd[{tag}]", msb, flop=c["pairs"] * 2 * H * 5 * d)


def bench_attn_queued(tag, B, N, H, lengths):
    """Short shapes: device time per C-ABI call with the host kept ahead of the GPU (the regime
    inside the captured train step), for the short-sequence kernels and, with GRB_NO_SHORT=1, the
    long-sequence kernels on the same inputs."""
    import os
    from mygenerativerecommenders_b200 import _lib
    c = attn_case(B, N, H, lengths)
    H, d = c["H"], c["d"]
    cache = GF.hstu_bucket_cache(c["off"], c["ts"], c["thr"], c["N"])
    for mode in ("short", "long"):
        if mode == "long":
            os.environ["GRB_NO_SHORT"] = "1"
        q, k, v = (c[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
        ts_w, pos_w = c["ts_w"].clone().requires_grad_(True), c["pos_w"].clone().requires_grad_(True)

        def step():
            out = GF.hstu_attention(q, k, v, c["off"], c["ts"], ts_w, pos_w, c["thr"], c["N"], H, d, d,
                                    bucket_cache=cache)
            torch.autograd.grad(out, (q, k, v, ts_w, pos_w), out)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _lib.profile_start()
        n_it = 10
        torch.cuda._sleep(20_000_000)
        for _ in range(n_it):
            step()
        prof = _lib.profile_stop()
        os.environ.pop("GRB_NO_SHORT", None)
        for name, (cnt, tot) in sorted(prof.items()):
            flop = None
            if name == "hstu_attn_fwd":
                flop = c["pairs"] * 2 * H * 2 * d
            if name == "hstu_attn_bwd":
                flop = c["pairs"] * 2 * H * 5 * d
            report(f"{name}[{tag}; {mode} kernels; queued]", tot / cnt, flop=flop, calls=cnt)


def bench_jagged():
    B, N, W = 128, 8192, 512
    lengths = torch.randint(1024, N + 1, (B,), generator=torch.Generator().manual_seed(0))
    off = ops.asynchronous_complete_cumsum(lengths.to(DEV))
    T = int(off[-1])
    jag = torch.randn(T, W, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.jagged_to_padded_dense(jag, off, N, 0.0))
    report("jagged_to_padded_dense[C5 128x8192x512 bf16]", ms, byts=(T + B * N) * W * 2)
    dense = ops.jagged_to_padded_dense(jag, off, N, 0.0)
    ms = timeit(lambda: ops.dense_to_jagged(dense, off, total=T))
    report("dense_to_jagged[C5 128x8192x512 bf16]", ms, byts=2 * T * W * 2)
    del dense
    x = torch.randn(T // 2, W, device=DEV, dtype=torch.bfloat16)
    u = torch.randn_like(x)
    ms = timeit(lambda: GF.layer_norm_gate(x, u, 1e-6))
    report("ln_gate_fwd[%dx512 bf16]" % x.shape[0], ms, byts=3 * x.numel() * 2)
    l64 = torch.randint(1, 200, (100_000,), device=DEV)
    ms = timeit(lambda: ops.asynchronous_complete_cumsum(l64), flush=False)
    report("complete_cumsum[B=100k]", ms, us=round(ms * 1e3, 2))


def bench_ssl():
    n, R, D, X = 14_000, 128, 256, 26_744
    g = torch.Generator(device=DEV).manual_seed(0)
    table = torch.nn.functional.normalize(torch.randn(X, D, device=DEV, generator=g), dim=-1).requires_grad_(True)
    q = torch.nn.functional.normalize(torch.randn(n, D, device=DEV, generator=g), dim=-1).requires_grad_(True)
    pos = torch.randint(0, X, (n,), device=DEV)
    idx = torch.randint(0, X, (n, R), device=DEV)
    p = table[pos].detach()
    byts = n * R * (D * 4 + 16) + 2 * n * D * 4 + n * (R + 1) * 4
    run = lambda: GF.sampled_softmax_rows(q, p, table, None, idx, None, pos, idx, False, 1e-6, 0.05)
    ms = timeit(run)
    report("sampled_softmax_fwd[C2 N'=14k R=128 D=256]", ms, byts=byts)
    rows = run()
    go = torch.ones_like(rows) / n
    msb = timeit(lambda: torch.autograd.grad(rows, (q, table), go, retain_graph=True), iters=5)
    report("sampled_softmax_bwd[C2]", msb, byts=byts + n * R * D * 4)


def bench_mips(tag, B, X, D, k, dtype):
    g = torch.Generator(device=DEV).manual_seed(0)
    items = torch.nn.functional.normalize(torch.randn(X, D, device=DEV, generator=g), dim=-1).to(dtype)
    q = torch.nn.functional.normalize(torch.randn(B, D, device=DEV, generator=g), dim=-1).to(dtype)
    es = items.element_size()
    ms = timeit(lambda: GF.mips_topk(q, items, None, k), iters=5, warmup=2)
    report(f"mips_topk[{tag}]", ms, flop=2.0 * B * X * D, byts=X * D * es + B * D * es + B * k * 12,
           qps=round(B / ms * 1e3, 1))


def bench_decode(tag, B, N, H, lengths, d=64):
    """Incremental path: the last position of every sequence against its K / V caches."""
    c = attn_case(B, N, H, lengths)
    off, lengths = c["off"], torch.as_tensor(lengths, dtype=torch.int64).to(DEV)
    kc = ops.jagged_to_padded_dense(c["k"], off, N, 0.0)
    pos = lengths - 1
    qn = c["q"][off[:-1] + pos]
    run = lambda: GF.hstu_attention_decode(qn, kc, c["v"], off, pos, c["ts"], c["ts_w"], c["pos_w"], c["thr"],
                                           N, H, d, d)
    byts = int(lengths.sum()) * H * 2 * d * 2 + B * H * 2 * d * 2 + int(lengths.sum()) * 8
    with torch.no_grad():
        report(f"hstu_attn_decode[{tag}]", timeit(run), byts=byts)


def timeit_queued(fn, n=20, rounds=5):
    """Per-call time of a short kernel with the host kept ahead of the GPU: park the GPU (~1.5 ms),
    queue n calls, time them as a block.  Inputs stay L2-warm, as they are inside the train step."""
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(rounds):
        torch.cuda._sleep(3_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best


def bench_silu():
    T, sizes = 14_336, [256, 256, 256, 256]
    x = torch.randn(T, 1024, device=DEV).to(torch.bfloat16).requires_grad_(True)
    byts = T * 1024 * 2
    tag = "C2 14336x1024 bf16, queued calls, L2-warm"
    with torch.no_grad():
        report(f"silu_fwd[{tag}]", timeit_queued(lambda: GF.silu_split(x, sizes)), byts=2 * byts)
        report(f"aten silu[{tag}] (library, for comparison)",
               timeit_queued(lambda: torch.nn.functional.silu(x)), byts=2 * byts)
    parts = GF.silu_split(x, sizes)
    gs = [torch.randn_like(p) for p in parts]
    report(f"silu_split_bwd[{tag}]",
           timeit_queued(lambda: torch.autograd.grad(parts, x, gs, retain_graph=True)), byts=3 * byts)
    ref = torch.split(torch.nn.functional.silu(x), sizes, dim=1)
    report(f"aten cat + silu_backward[{tag}] (library, for comparison)",
           timeit_queued(lambda: torch.autograd.grad(ref, x, gs, retain_graph=True)), byts=3 * byts)


def bench_adamw():
    from mygenerativerecommenders_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.randn(131_263, 256, device=DEV))
    p.grad = torch.randn_like(p) * 1e-3
    opt = FusedAdamW([p], lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
    ms = timeit(opt.step, flush=False)       # 941 MB per step: nothing survives in the 126 MB L2
    report("adamw_step[C2 item table 131263x256 fp32]", ms, byts=28 * p.numel())
    ref = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3, fused=True)
    ms = timeit(ref.step, flush=False)
    report("torch.optim.AdamW(fused=True)[same table] (library, for comparison)", ms, byts=28 * p.numel())


def main():
    which = set(sys.argv[1:])
    want = lambda n: not which or n in which
    if want("attn"):
        bench_attn("C5-slice 4x8192 H8 full", 4, 8192, 8, [8192] * 4)
        bench_attn("C5-jagged 8x U[1024,8192] H8", 8, 8192, 8,
                   torch.randint(1024, 8193, (8,), generator=torch.Generator().manual_seed(0)))
        bench_attn("C2 128x U[20,200] N211 H4", 128, 211, 4,
                   torch.randint(20, 201, (128,), generator=torch.Generator().manual_seed(0)))
    if "attnfwd" in which:
        bench_attn("prof 2x8192 H8 full", 2, 8192, 8, [8192] * 2, bwd=False)
    if "attnbwd" in which:
        bench_attn("prof 1x8192 H4 full", 1, 8192, 4, [8192], bwd=True)
    if "mipsshards" in which:    # what one rank of a 2- / 4- / 8-way sharded C4 index computes
        for g in (2, 4, 8):
            bench_mips(f"C4 shard 1/{g}: B4096 X{10_000_000 // g} D256 k200 bf16", 4096, 10_000_000 // g, 256,
                       200, torch.bfloat16)
    if "mipsshard8" in which:
        bench_mips("C4 shard 1/8: B4096 X1250000 D256 k200 bf16", 4096, 1_250_000, 256, 200, torch.bfloat16)
    if "mipsc3" in which:
        bench_mips("C3 B128 X700k D64 k261 bf16", 128, 700_000, 64, 261, torch.bfloat16)
    if "mipsshard4" in which:
        bench_mips("C4 shard 1/4: B4096 X2500000 D256 k200 bf16", 4096, 2_500_000, 256, 200, torch.bfloat16)
    if "mipsc4" in which:
        bench_mips("C4 B4096 X10M D256 k200 bf16", 4096, 10_000_000, 256, 200, torch.bfloat16)
    if "attnc2" in which:
        bench_attn("C2 128x U[20,200] N211 H4", 128, 211, 4,
                   torch.randint(20, 201, (128,), generator=torch.Generator().manual_seed(0)))
    if "attnc2q" in which:
        bench_attn_queued("C2 128x U[20,200] N211 H4", 128, 211, 4,
                          torch.randint(20, 201, (128,), generator=torch.Generator().manual_seed(0)))
        bench_attn_queued("C3 128x U[5,50] N61 H1", 128, 61, 1,
                          torch.randint(5, 51, (128,), generator=torch.Generator().manual_seed(0)))
    if "attnc5" in which:
        bench_attn("C5-slice 4x8192 H8 full", 4, 8192, 8, [8192] * 4)
    if "attnnobias" in which:
        c = attn_case(4, 8192, 8, [8192] * 4)
        H, d = c["H"], c["d"]
        q, k, v = (c[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
        out = GF.hstu_attention(q, k, v, c["off"], None, None, None, None, c["N"], H, d, d)
        report("hstu_attn_fwd[no-bias 4x8192 H8]", timeit(lambda: GF.hstu_attention(
            c["q"], c["k"], c["v"], c["off"], None, None, None, None, c["N"], H, d, d), flush=False),
            flop=c["pairs"] * 2 * H * 2 * d)
        go = torch.randn_like(out)
        report("hstu_attn_bwd[no-bias 4x8192 H8]", timeit(lambda: torch.autograd.grad(
            out, (q, k, v), go, retain_graph=True), iters=5, warmup=2, flush=False),
            flop=c["pairs"] * 2 * H * 5 * d)
    if want("decode"):
        g = torch.Generator().manual_seed(0)
        bench_decode("C2 128x U[20,200] N211 H4", 128, 211, 4, torch.randint(20, 201, (128,), generator=g))
        bench_decode("serving 4096x U[20,200] N211 H4", 4096, 211, 4, torch.randint(20, 201, (4096,), generator=g))
        bench_decode("C5 128x U[1024,8192] N8192 H8", 128, 8192, 8, torch.randint(1024, 8193, (128,), generator=g))
    if want("silu"):
        bench_silu()
    if want("adamw"):
        bench_adamw()
    if want("jagged"):
        bench_jagged()
    if want("ssl"):
        bench_ssl()
    if want("mips"):
        bench_mips("C3 B128 X700k D64 k261 bf16", 128, 700_000, 64, 261, torch.bfloat16)
        bench_mips("C3 B128 X700k D64 k261 fp32", 128, 700_000, 64, 261, torch.float32)
        bench_mips("C4 B4096 X10M D256 k200 bf16", 4096, 10_000_000, 256, 200, torch.bfloat16)
        bench_mips("C4-small-batch B128 X10M D256 k200 bf16", 128, 10_000_000, 256, 200, torch.bfloat16)


if __name__ == "__main__":
    main()
