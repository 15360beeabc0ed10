#!/usr/bin/env python
"""torch.profiler view of the C2 train step: host time per op, sync points, GPU busy time."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from mygenerativerecommenders_b200.pipeline import RetrievalModel, synthetic_batch, synthetic_item_ids

dev = torch.device("cuda")
cfg = bench.c2_config(True)
ids = synthetic_item_ids(26_744, cfg.num_items)
torch.manual_seed(42)
model = RetrievalModel(cfg, ids).to(dev).train()
import os
if os.environ.get("GRB_NO_GRAPHS") != "1":
    model.enable_step_graphs()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3, fused=True)
batches = [{k: v.to(dev) for k, v in synthetic_batch(cfg, ids, 128, seed=i).items()} for i in range(4)]
totals = [int(b["history_lengths"].sum()) for b in batches]

def step(i):
    loss = model.training_loss(batches[i % 4], totals[i % 4])
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss

for i in range(10):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    step(i)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 20 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=28, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
