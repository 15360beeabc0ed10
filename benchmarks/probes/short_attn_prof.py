"""A few forward + backward launches of the fused attention at the C2 shape (128 x U[20,200], 4
heads) for ncu: python benchmarks/probes/short_attn_prof.py [iters]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from kbench import attn_case  # noqa: E402
from mygenerativerecommenders_b200 import functional as GF  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lengths = torch.randint(20, 201, (128,), generator=torch.Generator().manual_seed(0))
c = attn_case(128, 211, 4, lengths)
cache = GF.hstu_bucket_cache(c["off"], c["ts"], c["thr"], c["N"])
q, k, v = (c[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
ts_w, pos_w = c["ts_w"].clone().requires_grad_(True), c["pos_w"].clone().requires_grad_(True)
for _ in range(iters):
    out = GF.hstu_attention(q, k, v, c["off"], c["ts"], ts_w, pos_w, c["thr"], 211, 4, 64, 64, bucket_cache=cache)
    torch.autograd.grad(out, (q, k, v, ts_w, pos_w), out)
torch.cuda.synchronize()
print("ok")
