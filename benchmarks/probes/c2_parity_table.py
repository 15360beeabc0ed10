"""Per-parameter gradient error of one C2-shape train step against the fp32 CPU oracle, for the
kernel variants selected by GRB_NO_SHORT / GRB_NO_PROJ_GEMM (developer probe)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "tests"))
import test_benchmark_shapes_gpu as T  # noqa: E402
from mygenerativerecommenders_b200.pipeline import synthetic_batch  # noqa: E402

ref_grads = None
for variant in ({}, {"GRB_NO_SHORT": "1"}, {"GRB_NO_PROJ_GEMM": "1"}, {"GRB_NO_SHORT": "1", "GRB_NO_PROJ_GEMM": "1"}):
    for k in ("GRB_NO_SHORT", "GRB_NO_PROJ_GEMM"):
        os.environ.pop(k, None)
    os.environ.update(variant)
    cfg, ids, m, ref = T._c2_pair()
    row = synthetic_batch(cfg, ids, 128, seed=1000)
    total = int(row["history_lengths"].sum())
    raw = torch.randint(0, 2 ** 40, (-(-total // 1024) * 1024, cfg.num_negatives), device="cuda",
                        generator=torch.Generator(device="cuda").manual_seed(1))
    T._inject_draws(m, raw)
    loss = m.training_loss({k: v.clone() for k, v in row.items()}, total_length=total)
    loss.backward()
    smp = m.negatives_sampler
    if ref_grads is None:
        picked = smp._cached_ids[(raw[:total] % smp._cached_count)].cpu()
        loss_ref = ref.training_loss(row, neg_draw=T._ref_draw(ref, row, picked))
        loss_ref.backward()
        ref_grads = {k.replace("|", "."): p.grad for k, p in ref.params.items()}
        ref_loss = loss_ref.item()
    worst = []
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        inf, l2 = T._rel(p.grad, ref_grads[k])
        worst.append((l2, inf, k))
    worst.sort(reverse=True)
    print(variant or "default", "loss rel", abs(loss.item() - ref_loss) / abs(ref_loss))
    for l2, inf, k in worst[:6]:
        print(f"   l2 {l2:.3e} inf {inf:.3e} {k}")
