"""Times the fused top-k in the one-query-block regime (C3 eval batch, C4 corpus at B=128): eager call and
the same call replayed as one CUDA graph.  GRB_MIPS_SMALL=0 selects the phased plan, GRB_MIPS_SMALL_STRIDE
the sample stride of the small-batch plan (both read once per process).
    python benchmarks/probes/mips_small_probe.py [c3] [c4]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from mygenerativerecommenders_b200 import functional as GF  # noqa: E402

DEV = "cuda"


def run(tag, B, X, D, k, n_inv):
    g = torch.Generator(device=DEV).manual_seed(0)
    items = torch.nn.functional.normalize(torch.randn(X, D, device=DEV, generator=g), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(B, D, device=DEV, generator=g), dim=-1).to(torch.bfloat16)
    inv = torch.randint(1, X, (B, n_inv), device=DEV, generator=g) if n_inv else None
    ids = torch.arange(1, X + 1, device=DEV, dtype=torch.int64) if os.environ.get("PROBE_IDS") == "1" else None
    graph = GF.MipsTopkGraph(B, items, ids, k, n_invalid=n_inv)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    res = {}
    for name, fn in (("eager", lambda: GF.mips_topk(q, items, ids, k, invalid_ids=inv)),
                     ("graph", lambda: graph(q, inv))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name + "_mean"] = sum(ts) / len(ts)
        ts.sort()
        res[name] = ts[len(ts) // 2]
    assert not graph.overflowed()
    byts = X * D * 2
    print(f"{tag} small={os.environ.get('GRB_MIPS_SMALL', '1')} stride={os.environ.get('GRB_MIPS_SMALL_STRIDE', 'auto')}: "
          f"eager {res['eager']:.4f} ms (mean {res['eager_mean']:.4f}), graph {res['graph']:.4f} ms (mean {res['graph_mean']:.4f}) "
          f"= {byts / res['graph'] / 1e6:.0f} GB/s, ids={'yes' if ids is not None else 'no'}", flush=True)


if __name__ == "__main__":
    which = set(sys.argv[1:]) or {"c3", "c4"}
    if "c3" in which:
        run("C3 128x700k x64 k200+61", 128, 700_000, 64, 200, 61)
    if "c4" in which:
        run("C4 128x10M x256 k200", 128, 10_000_000, 256, 200, 0)
