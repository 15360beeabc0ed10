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
from mygenerativerecommenders_b200.optim import FusedAdamW
opt = FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
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
SHAPES = os.environ.get("GRB_PROFILE_SHAPES") == "1"   # eager runs: attribute kernel time to ATen ops + shapes
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=SHAPES) as prof:
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
if SHAPES:
    print(prof.key_averages(group_by_input_shape=True).table(
        sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=50, max_shapes_column_width=70))
    # one step in launch order: every op that launched kernels itself, with its own GPU time
    evs = [e for e in prof.events() if e.self_device_time_total > 0 and e.device_type.name == "CPU"]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // 5
    print("# ops of the last profiled step, in launch order (self GPU us, name, input shapes)")
    for e in evs[-per:]:
        print(f"{e.self_device_time_total:9.1f}  {e.name[:48]:48s} {str(e.input_shapes)[:110]}")
    sys.exit(0)
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=28, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
