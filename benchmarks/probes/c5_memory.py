"""Where does the memory of a C5 train step go?  Prints allocated / peak bytes after each phase and the
largest live blocks (with the Python frames that allocated them) at the forward/backward boundary."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from mygenerativerecommenders_b200.optim import FusedAdamW  # noqa: E402
from mygenerativerecommenders_b200.pipeline import (RetrievalConfig, RetrievalModel, synthetic_batch,  # noqa: E402
                                                    synthetic_item_ids)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
cfg = RetrievalConfig(name="C5", num_items=131_262, max_sequence_length=8181, gr_output_length=10,
                      embedding_dim=512, num_blocks=8, num_heads=8, attention_dim=64, linear_dim=64, dropout=0.2,
                      sampler="local", num_negatives=128, temperature=0.05, top_k=200,
                      compute_dtype=torch.bfloat16)
ids = synthetic_item_ids(26_744, cfg.num_items)
torch.manual_seed(0)
model = RetrievalModel(cfg, ids).to(dev).train()
opt = FusedAdamW(model.parameters(), lr=1e-3)
row = {k: v.to(dev) for k, v in synthetic_batch(cfg, ids, B, seed=1, min_len=1024).items()}
tot = int(row["history_lengths"].sum())
gb = lambda x: f"{x / 2**30:.2f} GB"
print("tokens", tot, "model+state", gb(torch.cuda.memory_allocated()))
torch.cuda.memory._record_memory_history(max_entries=200000)
torch.cuda.reset_peak_memory_stats()
loss = model.training_loss(row, tot)
torch.cuda.synchronize()
print("after forward: allocated", gb(torch.cuda.memory_allocated()), "peak", gb(torch.cuda.max_memory_allocated()))
snap = torch.cuda.memory._snapshot()
blocks = []
for seg in snap["segments"]:
    for b in seg["blocks"]:
        if b["state"] == "active_allocated":
            fr = [f for f in b.get("frames", []) if "mygenerativerecommenders_b200" in f["filename"]]
            blocks.append((b["size"], [f"{Path(f['filename']).name}:{f['line']}" for f in fr[:3]]))
agg = {}
for size, fr in blocks:
    key = tuple(fr)
    a = agg.setdefault(key, [0, 0])
    a[0] += size
    a[1] += 1
for key, (size, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{gb(size):>10s} x{n:<4d} {' <- '.join(key)}")
torch.cuda.reset_peak_memory_stats()
loss.backward()
torch.cuda.synchronize()
print("after backward: allocated", gb(torch.cuda.memory_allocated()), "peak", gb(torch.cuda.max_memory_allocated()))
opt.step()
torch.cuda.synchronize()
print("after step: allocated", gb(torch.cuda.memory_allocated()), "peak", gb(torch.cuda.max_memory_allocated()))
