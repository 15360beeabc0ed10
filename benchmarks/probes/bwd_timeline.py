"""clock64 timeline of one CTA of the attention backward.  Needs a library built with
GRB_NVCC_EXTRA=-DGRB_BWD_TIMELINE python -m mygenerativerecommenders_b200.build"""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/benchmarks")
import os
import kbench
from mygenerativerecommenders_b200 import functional as GF
bias = sys.argv[1] == "bias"
if len(sys.argv) > 2 and sys.argv[2] == "c2":
    c = kbench.attn_case(128, 211, 4, torch.randint(150, 201, (128,), generator=torch.Generator().manual_seed(0)))
else:
    c = kbench.attn_case(4, 8192, 8, [8192] * 4)
H, d = c["H"], c["d"]
q, k, v = (c[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
if bias:
    cache = GF.hstu_bucket_cache(c["off"], c["ts"], c["thr"], c["N"])
    tw, pw = c["ts_w"].clone().requires_grad_(True), c["pos_w"].clone().requires_grad_(True)
    out = GF.hstu_attention(q, k, v, c["off"], c["ts"], tw, pw, c["thr"], c["N"], H, d, d, bucket_cache=cache)
    ins = (q, k, v, tw, pw)
else:
    out = GF.hstu_attention(q, k, v, c["off"], None, None, None, None, c["N"], H, d, d)
    ins = (q, k, v)
go = torch.randn_like(out)
os.environ["GRB_BWD_DEBUG"] = "0"
for _ in range(2): torch.autograd.grad(out, ins, go, retain_graph=True)
torch.cuda.synchronize()
os.environ["GRB_BWD_DEBUG"] = "8"
torch.autograd.grad(out, ins, go, retain_graph=True)
torch.cuda.synchronize()
