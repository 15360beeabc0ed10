"""Per-rank CUDA-event timeline of the data-parallel C2 train step: where the time
between the end of the local backward and the optimizer goes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 benchmarks/probes/dp_timeline.py [steps]

Events (all on the compute stream of each rank): step start | forward enqueued | local backward done |
touched table rows scaled + put to the peers, dense gradients copied into the symmetric buffer | barrier 0
passed (slowest rank arrived, puts landed) | peers' rows added + two-shot all-reduce kernel done | barrier 1
passed | optimizer done.  Prints, per rank, the mean of every span over the timed steps, and the tokens per
batch (the straggler effect: every barrier waits for the rank with the longest batch)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402
from mygenerativerecommenders_b200 import pipeline  # noqa: E402
from mygenerativerecommenders_b200.optim import FusedAdamW  # noqa: E402
from mygenerativerecommenders_b200.pipeline import RetrievalModel, synthetic_batch, synthetic_item_ids  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    world, rank, local = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
    dev = torch.device("cuda", local)
    cfg = bench.c2_config(bf16=True)
    ids = synthetic_item_ids(26_744, cfg.num_items)
    torch.manual_seed(42)
    model = RetrievalModel(cfg, ids).to(dev).train()
    n_batches = 8
    host = [synthetic_batch(cfg, ids, bench.PER_GPU_BATCH, seed=1000 * rank + i) for i in range(n_batches)]
    totals = [int(b["history_lengths"].sum()) for b in host]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    model.enable_step_graphs(row_granularity=1024, lazy=(world == 1))
    marks = {}

    def mark(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.setdefault(name, []).append(ev)

    step_mod = bench.TrainStep(model)
    reducer = None
    if world > 1:
        model.precapture_step_graphs(resident, totals)
        reducer = model.enable_peer_gradients()
        orig_barrier = reducer.sync.barrier

        def barrier(slot, device):
            mark(f"pre{slot}")
            orig_barrier(slot, device)
            mark(f"bar{slot}")
        reducer.sync.barrier = barrier
    opt = FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)

    def step(i, timed):
        if timed:
            mark("start")
        loss = step_mod(resident[i % n_batches], totals[i % n_batches])
        opt.zero_grad(set_to_none=True)
        if timed:
            mark("fwd")
        loss.backward()
        if timed:
            mark("bwd")
        if reducer is not None:
            reducer.reduce()
        if timed:
            mark("red")
        opt.step()
        if timed:
            mark("opt")

    for i in range(12):
        step(i, False)
    marks.clear()
    bench.barrier(world)
    for i in range(steps):
        step(i, True)
    torch.cuda.synchronize()

    def span(a, b):
        xs = [x.elapsed_time(y) for x, y in zip(marks[a], marks[b])]
        return sum(xs) / len(xs)
    out = {"rank": rank, "tokens_mean": sum(totals) / len(totals), "step_ms": span("start", "opt"),
           "fwd_ms": span("start", "fwd"), "bwd_ms": span("fwd", "bwd"), "reduce_ms": span("bwd", "red"),
           "opt_ms": span("red", "opt")}
    if world > 1:
        out.update({"put_rows_copy_dense_ms": span("bwd", "pre0"), "barrier0_ms": span("pre0", "bar0"),
                    "scatter_allreduce_ms": span("bar0", "pre1"), "barrier1_ms": span("pre1", "bar1")})
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, out)
    else:
        gathered = [out]
    if rank == 0:
        keys = [k for k in out if k != "rank"]
        print("rank " + " ".join(f"{k:>26s}" for k in keys))
        for o in gathered:
            print(f"{o['rank']:4d} " + " ".join(f"{o[k]:26.3f}" for k in keys))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
