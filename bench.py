#!/usr/bin/env python
"""bench.py — HSTU training sequences/s (+ top-k retrieval queries/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  Headline metric = BASELINE.json's: HSTU training sequences/s
on configs[1] (C2: ml-20m-shaped, N=211, 4 layers, D=256, 4 heads x 64, in-batch negatives
R=128, AdamW; bf16 activations / fp32 master weights), one full train step per "step"
(forward + backward + optimizer), data-parallel over ranks (weak scaling: 128 sequences per
GPU).  The same line carries a ``retrieval`` object with the second half of the metric: top-k
queries/s on C4 (10 M x 256 bf16 items sharded over the ranks, 4096 queries, k=200).

The line also carries, all measured in the same run: ``long_sequence`` (the long-sequence attention
kernels on a C5 slice and one full C5 train step), ``retrieval.small_batch`` / ``retrieval.c3`` (the
HBM-bound side of the top-k path), ``dropin_eager`` (the step as a reference user gets it: eager, no
graphs, torch AdamW) and ``loss_check`` (the benchmarked model's loss against the CPU oracle).

``--impl reference`` times the reference's CPU path on the host cores: the reference's own modules
from the staged copy under oracle/_ref (kind "reference"; oracle/stage_ref.py, run by build() where
/root/reference exists) or, without it, the restatement oracle/ref_step.py (kind "port"), on a
bounded sample of the same workload, and prints the same JSON shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from mygenerativerecommenders_b200.optim import FusedAdamW  # noqa: E402
from mygenerativerecommenders_b200.pipeline import (  # noqa: E402
    RetrievalConfig, synthetic_batch, synthetic_item_ids)

PER_GPU_BATCH = 128


def c2_config(bf16: bool = True) -> RetrievalConfig:
    return RetrievalConfig(
        name="C2 ml-20m-shaped", num_items=131_262, max_sequence_length=200, gr_output_length=10,
        embedding_dim=256, num_blocks=4, num_heads=4, attention_dim=64, linear_dim=64,
        dropout=0.2, sampler="inbatch", num_negatives=128, temperature=0.05, top_k=200,
        split_year_embedding=False, compute_dtype=torch.bfloat16 if bf16 else None)


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs_sustained": d.get("hbm_gbs_sustained", d["hbm_gbs"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "hbm_gbs_sustained": 6650.0, "source": "fallback"}


class ClockSampler:
    """SM clock / power / throttle reasons sampled every 10 ms while the timed region runs.

    In-process NVML (nvidia_ml_py): spawning `nvidia-smi -lms 200` next to the timed loop costs the
    training step ~40 % (its NVML session contends with kernel launches); the same counters read
    through one persistent handle do not.  Falls back to one nvidia-smi query if NVML is missing."""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples: list[tuple[float, float, int]] = []
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except (ValueError, IndexError):
                pass
        return self.idx

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
        return self

    def _poll(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            self._stop.wait(0.01)

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        return False

    def summary(self) -> dict:
        nv = self._nvml
        if nv is None or not self.samples:
            return self._smi_once()
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for s in self.samples:
            bits |= s[2]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, v in names.items() if bits & v)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self._max, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(s[1] for s in self.samples), "via": "nvml"}

    def _smi_once(self) -> dict:
        try:
            out = subprocess.run(
                ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                 "-i", str(self._physical_index())], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(x) for x in out.strip().split(",")[:2]]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "via": "nvidia-smi (after)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif n_gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run")
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world: int, dev) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------
# training arm (ours)
# --------------------------------------------------------------------------------------------
class TrainStep(torch.nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, row, total_length):
        return self.model.training_loss(row, total_length)


def attention_pairs(lengths: torch.Tensor) -> int:
    n = lengths.to(torch.int64)
    return int((n * (n + 1) // 2).sum().item())


def run_training(args, world, rank, local):
    from mygenerativerecommenders_b200 import _lib
    from mygenerativerecommenders_b200.pipeline import RetrievalModel
    dev = torch.device("cuda", local)
    cfg = c2_config(bf16=True)
    ids = synthetic_item_ids(26_744, cfg.num_items)
    torch.manual_seed(42)
    model = RetrievalModel(cfg, ids).to(dev).train()
    # the HSTU layer stack runs as captured CUDA graphs (one pair per padded row count); the eager
    # path is used for the per-kernel attribution pass below.  GRB_NO_GRAPHS=1 times the eager path.
    use_graphs = os.environ.get("GRB_NO_GRAPHS") != "1"
    n_batches = 8
    host = [synthetic_batch(cfg, ids, PER_GPU_BATCH, seed=1000 * rank + i) for i in range(n_batches)]
    totals = [int(b["history_lengths"].sum()) for b in host]
    if use_graphs:
        # the whole loss computation and its backward replay as one CUDA-graph pair per padded row
        # count; under DDP the graphs have to exist before the wrapper does (RetrievalModel.enable_step_graphs)
        model.enable_step_graphs(row_granularity=1024, lazy=(world == 1))
        if world > 1:
            model.precapture_step_graphs([{k: v.to(dev) for k, v in b.items()} for b in host], totals)
    step_mod = TrainStep(model)
    reducer = None
    if world > 1:
        if os.environ.get("GRB_DDP") != "1":
            # gradients are reduced over peer memory by this package's kernels: touched rows of the item
            # table + a two-shot all-reduce of the dense gradients, two stream barriers per step
            try:
                reducer = model.enable_peer_gradients()
            except Exception as e:   # no symmetric memory between these GPUs: DistributedDataParallel
                print(f"[bench] peer-memory gradient reduction unavailable ({e!r}); using DDP", file=sys.stderr)
        if reducer is None:
            if os.environ.get("GRB_NO_P2P") != "1":
                try:
                    ignore = ["model." + n for n in model.enable_peer_table_grads()]
                    torch.nn.parallel.DistributedDataParallel._set_params_and_buffers_to_ignore_for_model(
                        step_mod, ignore)
                except Exception as e:
                    print(f"[bench] peer-memory table gradients unavailable ({e!r}); dense all-reduce by DDP",
                          file=sys.stderr)
            step_mod = torch.nn.parallel.DistributedDataParallel(
                step_mod, device_ids=[local], gradient_as_bucket_view=True, broadcast_buffers=False)
    if os.environ.get("GRB_TORCH_ADAMW") == "1":     # A/B switch: the library optimizer
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3, fused=True)
    else:
        opt = FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def step(row, total):
        loss = step_mod(row, total)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if reducer is not None:
            reducer.reduce()
        opt.step()
        return loss

    # untimed setup (not part of --warmup): cuBLAS heuristics, allocator growth, DDP bucket
    # rebuild and NCCL channel setup all happen in the first handful of steps
    for i in range(8):
        step(resident[i % n_batches], totals[i % n_batches])
    for i in range(args.warmup):
        step(resident[i % n_batches], totals[i % n_batches])
    barrier(world)

    # ---- timed region: inputs resident in HBM -------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier(world)
        e0.record()
        for i in range(args.steps):
            step(resident[i % n_batches], totals[i % n_batches])
        e1.record()
        barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)

    # ---- end to end: pinned host batch -> device each step, loss read back each step ---------
    h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
    barrier(world)
    # The loss of every step is copied to pinned host memory as part of the step and read by the
    # host one step later (what a logging callback does): a blocking .item() right after launching
    # the step would only measure how long the host takes to enqueue the next one.
    last = 0.0
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        row = {k: v.to(dev, non_blocking=True) for k, v in pinned[i % n_batches].items()}
        loss = step(row, totals[i % n_batches])
        loss_host[i % 2].copy_(loss.detach(), non_blocking=True)       # D2H read of the loss
        copied[i % 2].record()
        if i > 0:
            copied[(i - 1) % 2].synchronize()
            last = float(loss_host[(i - 1) % 2])
    copied[(args.steps - 1) % 2].synchronize()
    last = float(loss_host[(args.steps - 1) % 2])
    e3.record()
    barrier(world)
    ms_e2e = max_over_ranks(e2.elapsed_time(e3), world, dev)

    # ---- per-kernel attribution: the same steps run eagerly with CUDA events around every
    # C-ABI call (a graph replay offers no place to put them); also counts our launches per step
    n_attr = min(args.steps, 10)
    if use_graphs:
        model.disable_cuda_graphs()
    for i in range(2):
        step(resident[i % n_batches], totals[i % n_batches])
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    _lib.profile_start()
    for i in range(n_attr):
        # the eager step is host-bound: without a head start the interval between the two events of
        # a kernel would include the GPU waiting for the host to launch it.  Park the GPU for ~3 ms
        # so that the host is ahead and the kernels run back to back.
        torch.cuda._sleep(6_000_000)
        step(resident[i % n_batches], totals[i % n_batches])
    prof = _lib.profile_stop()
    launches = (_lib.launch_count() - launches0) * args.steps // n_attr
    barrier(world)

    # ---- the drop-in path as a reference user gets it (world == 1): same modules, eager (no CUDA
    # graphs), no host-provided row count (so the jagged ops read their sizes back from the device,
    # as every call site of the reference does), torch.optim.AdamW instead of the fused optimizer
    dropin = None
    if world == 1:
        opt_t = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)

        def step_t(row):
            loss = model.training_loss(row)
            opt_t.zero_grad(set_to_none=True)
            loss.backward()
            opt_t.step()
            return loss
        n_d = max(5, min(args.steps, 40))
        for i in range(3):
            step_t(resident[i % n_batches])
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for i in range(n_d):
            row = {k: v.to(dev, non_blocking=True) for k, v in pinned[i % n_batches].items()}
            last_d = float(step_t(row).detach())             # blocking loss read, as a logger would
        d1.record()
        torch.cuda.synchronize()
        ms_d = d0.elapsed_time(d1)
        dropin = {"value": PER_GPU_BATCH * n_d / (ms_d / 1e3), "unit": "sequences/s", "ms_per_step": ms_d / n_d,
                  "steps": n_d, "our_launches_per_step": (_lib.launch_count() - l0) / n_d,
                  "what": "end to end (pinned H2D batch, blocking loss read), eager modules behind the reference's "
                          "module boundary, no CUDA graphs, no host-provided total_length, torch.optim.AdamW"}
        del opt_t

    # ---- parity probe on the benchmarked model: eval mode (no dropout), fixed negative draws; the CPU
    # side of the comparison runs in the cpu_baseline leg (oracle/ref_step.py on the same weights)
    probe = None
    if world == 1:
        model.eval()
        smp = model.negatives_sampler
        row0 = host[0]
        tot0 = totals[0]
        raw = torch.randint(0, 2 ** 40, (tot0, cfg.num_negatives), device=dev,
                            generator=torch.Generator(device=dev).manual_seed(1))
        smp._draw = lambda positive_ids, n: raw[: positive_ids.size(0)] % (
            smp._cached_count if smp._cached_count is not None else smp._cached_ids.size(0))
        with torch.no_grad():
            l_gpu = float(model.training_loss({k: v.to(dev) for k, v in row0.items()}, total_length=tot0))
        picked = smp._cached_ids[(raw % smp._cached_count)].cpu()
        del smp.__dict__["_draw"]              # back to the class's draw (and its fused kernel)
        model.train()
        probe = {"row": row0, "picked": picked, "loss_gpu": l_gpu}

    seqs = PER_GPU_BATCH * world * args.steps
    out = {
        "value": seqs / (ms / 1e3), "ms_per_step": ms / args.steps,
        "e2e": {"value": seqs / (ms_e2e / 1e3), "unit": "sequences/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "d2h": "loss of every step copied to pinned host memory, consumed by the host one step later"},
        "gpu_launches": int(launches), "clocks": clk.summary(), "final_loss": last,
        "dropin_eager": dropin, "_probe": probe,
    }
    # ---- roofline of the dominant hand-written kernel, from the live per-kernel events -------
    pk = peaks()
    lengths_used = [host[i % n_batches]["history_lengths"] for i in range(n_attr)]
    pairs = sum(attention_pairs(l) for l in lengths_used) * cfg.num_blocks
    rows = sum(int(l.sum()) for l in lengths_used)          # supervised positions N' per step, summed
    H, dqk, dv, D, R = cfg.num_heads, cfg.attention_dim, cfg.linear_dim, cfg.embedding_dim, cfg.num_negatives
    ssl_bytes = rows * R * (D * 4 + 16) + 2 * rows * D * 4 + rows * (R + 1) * 4
    work = {  # name -> (bound, algorithmic work summed over the attribution pass)
        "hstu_attn_fwd": ("tensor", pairs * 2 * H * (dqk + dv)),
        "hstu_attn_bwd": ("tensor", pairs * 2 * H * (3 * dqk + 2 * dv)),
        "sampled_softmax_fwd": ("hbm", ssl_bytes),
        # bf16 backward (csrc/ssl_bwd_csr.cu): two gathers of bf16 rows (cache rows for dq, q rows for the
        # cache gradient) + index / coefficient / pair arrays, + q, p, dq, dp rows
        "sampled_softmax_bwd": ("hbm", rows * R * (2 * D * 2 + 28) + 4 * rows * D * 4),
        # 4 loads + 3 stores of fp32 per parameter element (include/grb200.h grb_adamw_step)
        "adamw_step": ("hbm", 28 * sum(p.numel() for p in model.parameters()) * n_attr),
    }
    # the six projection GEMMs of a layer (csrc/proj_gemm.cu), 2 M N K each, over the padded row bucket
    t_pads = sum(-(-int(l.sum()) // 1024) * 1024 for l in lengths_used)
    n_uvqk, n_o = 2 * H * dv + 2 * H * dqk, H * dv
    for kind in ("fwd", "dgrad", "wgrad"):
        work[f"proj_gemm_uvqk_{kind}"] = ("tensor", cfg.num_blocks * 2 * t_pads * D * n_uvqk)
        work[f"proj_gemm_o_{kind}"] = ("tensor", cfg.num_blocks * 2 * t_pads * n_o * D)
    step_ms = ms / args.steps
    kern = {}
    for name, (n, tot_ms) in prof.items():
        kern[name] = {"calls_per_step": n / n_attr, "ms_per_step": round(tot_ms / n_attr, 4),
                      "share_of_step": round(tot_ms / n_attr / step_ms, 4) if step_ms > 0 else None}
    out["kernels"] = kern
    out["kernels_timed_in"] = (f"eager attribution pass of {n_attr} steps after the timed region "
                               "(CUDA events around each C-ABI call); shares are against the timed step")
    def line(name):
        n, tot_ms = prof[name]
        bound, amount = work[name]
        if bound == "tensor":
            achieved, peak, unit = amount / (tot_ms / 1e3) / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s"
        else:
            achieved, peak, unit = amount / (tot_ms / 1e3) / 1e9, pk["hbm_gbs_sustained"], "GB/s"
        # `traffic` is by definition a profiler figure (dram bytes of one `ncu --set full` capture of this
        # kernel at this workload): it cannot be measured inside a timed run, so it carries its source
        traffic, traffic_src = None, None
        tf = ROOT / "profiles" / "r2_traffic.json"
        if tf.exists():
            ent = json.loads(tf.read_text()).get(name)
            if ent and ent.get("matches_bench_workload"):
                traffic, traffic_src = ent["dram_bytes_per_launch"], f"profiles/r2_traffic.json <- {ent.get('capture')}"
        r = {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
             "peak_source": pk["source"] + " (sustained figure: kernel timed inside a long step)",
             "algorithmic_work_per_launch": amount / max(n, 1), "avg_launch_ms": tot_ms / max(n, 1)}
        if name.startswith("hstu_attn"):
            r["note"] = ("C2 sequences are <= 211 tokens: the short-sequence kernels (csrc/hstu_attn_short.cu, "
                         "whole sequence per CTA, two CTAs per SM) run here; ~45 % of the executed tile area is "
                         "causal / ragged padding.  The long-sequence kernels are timed live in `long_sequence`.")
        if name.startswith("sampled_softmax"):
            # the gather source here is the 11 MB in-batch cache, which stays in L2, so
            # the algorithmic gather bytes are served above the HBM peak; the L2 cap is the real bound
            l2_cap = 6300.0 * clk.summary().get("sm_mhz", 1965.0) * 1e6 / 1e9     # B/clk (B300 guide) x SM clock
            r["note"] = ("the gather sources (in-batch cache, output embeddings) are L2-resident: frac > 1 against "
                         "HBM is expected, l2_frac is against ~6300 B/clk of L2 bandwidth; the backward is five "
                         "launches timed together (prep, rows, scan, scatter, cols)")
            r["l2_frac"] = achieved / l2_cap
        return r

    present = [k for k in work if k in prof]
    if present:
        dom = max(present, key=lambda k: prof[k][1])       # largest share of the step, whatever it is
        out["roofline"] = line(dom)
        out["roofline_others"] = [line(k) for k in present if k != dom]
    return out, cfg, ids, model


# --------------------------------------------------------------------------------------------
# long sequences (C5): the configuration on which "attention >= 60 % of the tensor pipe" is testable
# --------------------------------------------------------------------------------------------
def _event_ms(fn, iters: int, warmup: int) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_long_sequence(local: int, world: int = 1, rank: int = 0) -> dict:
    """Measured in this run (nothing read from profiles/): (1) the long-sequence attention kernels on a
    C5 slice, 4 sequences x 8192 tokens, H=8, d=64, forward and backward timed alone (burst peak);
    (2) one full C5 training step (BASELINE.json configs[4]: N=8192, 8 layers, D=512, 8 heads x 64,
    bf16, local negatives R=128, AdamW), jagged lengths U[1024, 8181], at the largest batch <= 128 per GPU
    that fits; data parallel over the ranks when launched with torchrun (gradients, the dense 269 MB item
    table included, reduced by PeerGradients' two-shot all-reduce kernel; max over ranks of the event time)."""
    from mygenerativerecommenders_b200 import _lib
    from mygenerativerecommenders_b200 import functional as GF
    from mygenerativerecommenders_b200 import hstu
    from mygenerativerecommenders_b200.pipeline import RetrievalModel
    dev = torch.device("cuda", local)
    pk = peaks()
    out = {}
    # ---- (1) attention slice ---------------------------------------------------------------
    if world == 1:
        B, N, H, d = 4, 8192, 8, 64
        lengths = torch.full((B,), N, dtype=torch.int64)
        off = torch.zeros(B + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(lengths, 0)
        T = int(off[-1])
        g = torch.Generator(device=dev).manual_seed(0)
        mk = lambda: (torch.randn(T, H * d, device=dev, generator=g) * 0.5).to(torch.bfloat16).requires_grad_(True)
        q, k, v = mk(), mk(), mk()
        ts = (978_300_000 + torch.cumsum(torch.randint(1, 5000, (B, N)), 1)).to(dev)
        ts_w = (torch.randn(129, device=dev) * 0.02).requires_grad_(True)
        pos_w = (torch.randn(2 * N - 1, device=dev) * 0.02).requires_grad_(True)
        thr = hstu.tabulate_bucket_thresholds(hstu._default_bucketization, 128).to(dev)
        offd = off.to(dev)
        cache = GF.hstu_bucket_cache(offd, ts, thr, N)
        pairs = int((lengths * (lengths + 1) // 2).sum())
        fwd = lambda: GF.hstu_attention(q, k, v, offd, ts, ts_w, pos_w, thr, N, H, d, d, bucket_cache=cache)
        with torch.no_grad():
            ms_f = _event_ms(fwd, 10, 3)
        o = fwd()
        go = torch.randn_like(o)
        bwd = lambda: torch.autograd.grad(o, (q, k, v, ts_w, pos_w), go, retain_graph=True)
        ms_b = _event_ms(bwd, 5, 2)
        f_f, f_b = pairs * 2 * H * 2 * d, pairs * 2 * H * 5 * d
        out["attention_slice"] = {
            "shape": "4 x 8192 tokens, H=8, d=64, bf16, relative time+position bias (C5 slice)",
            "fwd_ms": ms_f, "fwd_tflops": f_f / ms_f / 1e9, "fwd_frac": f_f / ms_f / 1e9 / pk["bf16_tflops"],
            "bwd_ms": ms_b, "bwd_tflops": f_b / ms_b / 1e9, "bwd_frac": f_b / ms_b / 1e9 / pk["bf16_tflops"],
            "peak": pk["bf16_tflops"], "peak_source": pk["source"] + " (burst figure: kernels timed alone)",
            "measured": "live, CUDA events, 10 / 5 launches after warm-up"}
        del q, k, v, o, go, cache
        torch.cuda.empty_cache()
    # ---- (2) full C5 step ------------------------------------------------------------------
    cfg = RetrievalConfig(
        name="C5 long-sequence", num_items=131_262, max_sequence_length=8181, gr_output_length=10,
        embedding_dim=512, num_blocks=8, num_heads=8, attention_dim=64, linear_dim=64, dropout=0.2,
        sampler="local", num_negatives=128, temperature=0.05, top_k=200, split_year_embedding=False,
        compute_dtype=torch.bfloat16)
    ids = synthetic_item_ids(26_744, cfg.num_items)
    step_info = None
    for Bc in ((128, 64, 32, 16) if world == 1 else (128,)):
        model = opt = rows = reducer = None
        try:
            torch.manual_seed(42)
            model = RetrievalModel(cfg, ids).to(dev).train()
            if world > 1:
                reducer = model.enable_peer_gradients()
            opt = FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
            rows = [{k_: v_.to(dev) for k_, v_ in
                     synthetic_batch(cfg, ids, Bc, seed=500 + 10 * rank + i, min_len=1024).items()}
                    for i in range(2)]
            totals = [int(r["history_lengths"].sum()) for r in rows]
            torch.cuda.reset_peak_memory_stats(dev)

            def step(i):
                loss = model.training_loss(rows[i % 2], totals[i % 2])
                opt.zero_grad(set_to_none=True)
                loss.backward()
                if reducer is not None:
                    reducer.reduce()
                opt.step()
                return loss
            step(0)
            barrier(world)
            _lib.profile_start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_it = 3
            e0.record()
            for i in range(n_it):
                loss = step(i + 1)
            e1.record()
            barrier(world)
            prof = _lib.profile_stop()
            ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / n_it
            tok_all = float(sum(totals)) / 2
            if world > 1:
                import torch.distributed as dist
                tt = torch.tensor([tok_all], device=dev, dtype=torch.float64)
                dist.all_reduce(tt)
                tok_all = float(tt.item())
            lens = [rows[(i + 1) % 2]["history_lengths"] for i in range(n_it)]
            pr = sum(attention_pairs(l.cpu() + 1) for l in lens) * cfg.num_blocks
            Hh, dd = cfg.num_heads, cfg.attention_dim
            att = {}
            for name, mult in (("hstu_attn_fwd", 2), ("hstu_attn_bwd", 5)):
                if name in prof:
                    n_, tot_ = prof[name]
                    fl = pr * 2 * Hh * mult * dd
                    att[name] = {"ms_per_step": tot_ / n_it, "tflops": fl / tot_ / 1e9,
                                 "frac": fl / tot_ / 1e9 / pk["bf16_tflops_sustained"]}
            step_info = {
                "workload": f"C5 (BASELINE.json configs[4]) on {world} GPU(s), data parallel: N=8192, 8 HSTU layers, "
                            "D=512, 8 heads x 64, bf16 / fp32 master, local negatives R=128, dropout 0.2, AdamW; "
                            "jagged lengths U[1024, 8181]; fwd+bwd+optimizer, eager (no graphs)",
                "batch_per_gpu": Bc, "n_gpus": world, "tokens_per_step": tok_all, "ms_per_step": ms,
                "sequences_per_s": Bc * world / (ms / 1e3), "tokens_per_s": tok_all / (ms / 1e3),
                "attention_figures": "this rank's kernels",
                "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                "attention": att, "attention_peak": pk["bf16_tflops_sustained"],
                "attention_share_of_step": sum(a["ms_per_step"] for a in att.values()) / ms if att else None,
                "loss": float(loss.detach()), "steps_timed": n_it}
            break
        except torch.OutOfMemoryError:
            step_info = {"error": f"out of memory at batch {Bc}"}
        finally:
            del model, opt, rows, reducer
            torch.cuda.empty_cache()
    out["train_step"] = step_info
    return out


# --------------------------------------------------------------------------------------------
# retrieval arm (ours): C4, corpus sharded over the ranks
# --------------------------------------------------------------------------------------------
def run_retrieval(args, world, rank, local, X_total=10_000_000, B=4096, D=256, k=200):
    from mygenerativerecommenders_b200 import _lib
    from mygenerativerecommenders_b200 import functional as GF
    from mygenerativerecommenders_b200.candidate_index import merge_sharded_topk
    dev = torch.device("cuda", local)
    per = -(-X_total // world)
    lo, hi = min(rank * per, X_total), min((rank + 1) * per, X_total)
    g = torch.Generator(device=dev).manual_seed(rank)
    items = torch.nn.functional.normalize(
        torch.randn(hi - lo, D, device=dev, generator=g), dim=-1).to(torch.bfloat16)
    item_ids = torch.arange(lo + 1, hi + 1, device=dev, dtype=torch.int64)
    gq = torch.Generator(device=dev).manual_seed(1234)
    n_q = 4
    queries = [torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=gq), dim=-1)
               .to(torch.bfloat16) for _ in range(n_q)]
    pinned_q = [q.cpu().pin_memory() for q in queries]

    def finish(call):
        s, i = call.result()            # waits for this call's overflow flag only; exact re-run if set
        if world > 1:
            s, i = merge_sharded_topk(s, i, k, world, k_locals=[k] * world)   # peer-memory exchange
        return s, i

    def run_resident(n):
        # one call in flight, as in the end-to-end loop below: the host never waits for the device
        pending = None
        for i in range(n):
            call = GF.mips_topk_async(queries[i % n_q], items, item_ids, k)
            if pending is not None:
                finish(pending)
            pending = call
        finish(pending)

    steps = max(2, min(args.steps, 10))
    run_resident(2)
    barrier(world)
    _lib.profile_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_resident(steps)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    prof = _lib.profile_stop()
    # results land in pinned host buffers (two, alternating) and are consumed one step later, so the
    # copy of step i overlaps the host work of step i + 1 instead of blocking it
    ids_host = [torch.empty((B, k), dtype=torch.int64).pin_memory() for _ in range(2)]
    checksum = 0

    def collect(call, i):
        s_, ids_ = call.result()                                # exact: re-runs if a row overflowed
        if world > 1:
            s_, ids_ = merge_sharded_topk(s_, ids_, k, world, k_locals=[k] * world)
        ids_host[i % 2].copy_(ids_, non_blocking=False)         # D2H of the result, read by the host
        return int(ids_host[i % 2][0, 0])
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    # one call in flight: step i is enqueued (its overflow flag travels to pinned memory behind the
    # ids), then the host collects step i - 1 -- the flag check of MipsTopkCall.result() and the ids
    pending = None
    for i in range(steps):
        q = pinned_q[i % n_q].to(dev, non_blocking=True)
        call = GF.mips_topk_async(q, items, item_ids, k)
        if pending is not None:
            checksum += collect(*pending)
        pending = (call, i)
    checksum += collect(*pending)
    e3.record()
    barrier(world)
    ms_e2e = max_over_ranks(e2.elapsed_time(e3), world, dev)
    pk = peaks()
    extras = {}
    if world == 1:
        extras = retrieval_hbm_regime(dev, items, item_ids, pk)
    n_calls, k_ms = prof.get("mips_topk", (steps, ms))
    flop = 2.0 * B * (hi - lo) * D
    byts = (hi - lo) * D * 2 + B * D * 2 + B * k * 12
    t_kernel = k_ms / max(n_calls, 1) / 1e3
    return {
        "metric": "topk_retrieval_queries_per_s", "unit": "queries/s",
        "value": B * steps / (ms / 1e3), "ms_per_step": ms / steps, "steps": steps,
        "e2e": {"value": B * steps / (ms_e2e / 1e3), "unit": "queries/s",
                "h2d_bytes_per_step": B * D * 2, "d2h_bytes_per_step": B * k * 8,
                "d2h": "top-k ids of every step copied to pinned host memory, consumed one step later"},
        "config": {"workload": f"C4 retrieval-only MIPS top-k={k}: {X_total} x {D} bf16 items "
                               f"sharded over {world} GPU(s), {B} queries/batch, exact, "
                               "ties->lowest id; table (5.1 GB total) >> L2",
                   "items_per_gpu": hi - lo},
        "roofline": {"kernel": "grb_mips_topk (score GEMM + select)", "bound": "tensor",
                     "achieved": flop / t_kernel / 1e12, "peak": pk["bf16_tflops"],
                     "unit": "TFLOP/s", "frac": flop / t_kernel / 1e12 / pk["bf16_tflops"],
                     "hbm_frac": byts / t_kernel / 1e9 / pk["hbm_gbs"], "traffic": None,
                     "peak_source": pk["source"]},
        **extras,
    }


def retrieval_hbm_regime(dev, items, item_ids, pk) -> dict:
    """The HBM-bound side of the top-k path (B below ~250 queries), measured live:
    (1) the C4 corpus at B=128; (2) C3 (BASELINE.json configs[2]): 128 queries x 700 000 x 64 bf16,
    k=200 with the users' 61 past ids filtered inside the selection (k' = 261 never leaves the
    kernel); L2 flushed before every timed call (the 90 MB C3 table would otherwise stay resident)."""
    from mygenerativerecommenders_b200 import functional as GF
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    # These calls are 0.1 - 1 ms long and, at C3, instruction-bound: they follow the SM clock.  The timed C4
    # loop just before draws > 600 W; measured right behind it the same C3 graph takes 0.115 ms instead of
    # the 0.087 ms it takes in a process of its own (benchmarks/probes/mips_small_probe.py, same box).  Let
    # the board settle, and sample the clocks over this region so that the line says under which ones it ran.
    torch.cuda.synchronize(dev)
    time.sleep(2.0)

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    def launches_of(fn):
        from mygenerativerecommenders_b200 import _lib
        n0 = _lib.launch_count()
        fn()
        return _lib.launch_count() - n0

    g = torch.Generator(device=dev).manual_seed(7)
    X, D, k = items.shape[0], items.shape[1], 200
    q = torch.nn.functional.normalize(torch.randn(128, D, device=dev, generator=g), dim=-1).to(torch.bfloat16)
    sampler = ClockSampler(dev.index or 0)
    sampler.__enter__()
    ms = timed(lambda: GF.mips_topk(q, items, item_ids, k))
    byts = X * D * 2 + 128 * D * 2 + 128 * k * 12
    gr = GF.MipsTopkGraph(128, items, item_ids, k)
    ms_g = timed(lambda: gr(q))
    assert not gr.overflowed()
    out["small_batch"] = {"workload": f"C4 corpus ({X} x {D} bf16), 128 queries, k={k}", "ms": ms,
                          "graph_ms": ms_g, "graph_hbm_frac": byts / ms_g / 1e6 / pk["hbm_gbs"],
                          "queries_per_s": 128 / (ms / 1e3), "gbs": byts / ms / 1e6,
                          "hbm_frac": byts / ms / 1e6 / pk["hbm_gbs"],
                          "tensor_frac": 2.0 * 128 * X * D / ms / 1e9 / pk["bf16_tflops"],
                          "launches_per_call": launches_of(lambda: GF.mips_topk(q, items, item_ids, k)),
                          "plan": "4 launches = the one-query-block plan of csrc/mips_small.cu (group-maxima sample, "
                                  "threshold, private sub-lists, select); 9+ = the phased plan"}
    X3, D3, n_inv = 700_000, 64, 61
    items3 = torch.nn.functional.normalize(torch.randn(X3, D3, device=dev, generator=g), dim=-1).to(torch.bfloat16)
    ids3 = torch.arange(1, X3 + 1, device=dev, dtype=torch.int64)
    q3 = torch.nn.functional.normalize(torch.randn(128, D3, device=dev, generator=g), dim=-1).to(torch.bfloat16)
    inv3 = torch.randint(1, X3 + 1, (128, n_inv), device=dev, generator=g)
    ms3 = timed(lambda: GF.mips_topk(q3, items3, ids3, k, invalid_ids=inv3))
    byts3 = X3 * D3 * 2 + 128 * D3 * 2 + 128 * k * 12 + 128 * n_inv * 8
    gr3 = GF.MipsTopkGraph(128, items3, ids3, k, n_invalid=n_inv)
    ms3_g = timed(lambda: gr3(q3, inv3))
    assert not gr3.overflowed()
    assert torch.equal(gr3(q3, inv3)[1], GF.mips_topk(q3, items3, ids3, k, invalid_ids=inv3)[1])
    out["c3"] = {"workload": f"C3 eval batch: 128 queries x {X3} x {D3} bf16, k={k}, {n_inv} invalid ids per "
                             "query filtered inside the selection", "ms": ms3,
                 "graph_ms": ms3_g, "graph_hbm_frac": byts3 / ms3_g / 1e6 / pk["hbm_gbs"],
                 "graph": "the same call replayed as one CUDA graph (GF.MipsTopkGraph: static buffers, no Python "
                          "between the launches)",
                 "queries_per_s": 128 / (ms3 / 1e3), "gbs": byts3 / ms3 / 1e6,
                 "hbm_frac": byts3 / ms3 / 1e6 / pk["hbm_gbs"], "hbm_peak": pk["hbm_gbs"],
                 "launches_per_call": launches_of(lambda: GF.mips_topk(q3, items3, ids3, k, invalid_ids=inv3)),
                 "includes": "host time of the call (workspace lookup, launches, overflow-flag read)"}
    sampler.__exit__()
    out["clocks"] = dict(sampler.summary(), what="SM clocks over the small_batch and c3 measurements (after a 2 s idle "
                                               "behind the timed C4 loop)")
    return out


# --------------------------------------------------------------------------------------------
# CPU baseline (the reference's CPU path, oracle port)
# --------------------------------------------------------------------------------------------
def cpu_train_baseline(cfg, ids, sample_batch: int, steps: int, warmup: int, state_dict=None):
    """The reference's CPU path on the host cores.  With the staged package (oracle/_ref, written by
    oracle/stage_ref.py where /root/reference exists) the reference's OWN modules run, unmodified, with
    their stock Python fallbacks for the jagged ops -- that is `trainer=cpu` (kind "reference").  Without
    it: oracle/ref_step.py, the restatement (kind "port", ~4x faster than the reference itself)."""
    from oracle import ref_verbatim
    from mygenerativerecommenders_b200.pipeline import RetrievalModel
    torch.set_num_threads(os.cpu_count() or 1)
    fp32 = RetrievalConfig(**{**cfg.__dict__, "compute_dtype": None})
    if ref_verbatim.available():
        torch.manual_seed(42)
        ref = ref_verbatim.ReferenceRetrieval(fp32, ids).train()
        kind = "reference"
        how = ("the reference's own modules (staged copy of generative_recommenders_pl, unmodified: HSTU with "
               "the stock Python fallbacks of the jagged ops, in-batch sampler, SampledSoftmaxLoss) driven by "
               "the restated Retrieval.training_step")
    else:
        from oracle.ref_step import RefRetrieval
        if state_dict is None:
            torch.manual_seed(42)
            state_dict = RetrievalModel(fp32, ids).state_dict()
        ref = RefRetrieval.from_state_dict(fp32, ids, state_dict).train()
        kind = "port"
        how = ("oracle/ref_step.py: the reference formulation restated (padded attention, python-loop jagged "
               "ops, materialised negatives)")
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
    batches = [synthetic_batch(fp32, ids, sample_batch, seed=77 + i) for i in range(2)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = ref.training_loss(batches[i % 2])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": sample_batch * len(times) / total, "unit": "sequences/s",
            "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{len(times)} full train steps (fwd+bwd+AdamW) of the C2 model at batch {sample_batch} "
                      f"(same shapes, fp32): {how}; {total:.1f} s of CPU work",
            "ms_per_step": 1e3 * total / len(times)}


def cpu_loss_check(cfg, ids, state_dict, probe) -> dict:
    """The benchmarked model's loss on batch 0 (eval mode, the negative draws the GPU made) recomputed
    by the CPU oracle (oracle/ref_step.py) from the same weights: DESIGN.md section 2 tolerance 1e-3."""
    from oracle import reference_port as O
    from oracle.ref_step import RefRetrieval
    fp32 = RetrievalConfig(**{**cfg.__dict__, "compute_dtype": None})
    ref = RefRetrieval.from_state_dict(fp32, ids, state_dict).eval()
    row = probe["row"]
    lengths, pids, _ = ref.features(row)
    pids = pids.scatter(1, lengths.view(-1, 1), row["target_ids"].view(-1, 1))
    flat = pids.reshape(-1)
    with torch.no_grad():
        cid, _ = O.inbatch_process(flat, flat != 0, ref.item_emb(flat), ref.cfg.l2_eps, True)
        order = torch.argsort(cid)
        pos = torch.searchsorted(cid[order], probe["picked"])
        l_cpu = float(ref.training_loss(row, neg_draw=order[pos]))
    rel = abs(probe["loss_gpu"] - l_cpu) / abs(l_cpu)
    out = {"loss_gpu_bf16": probe["loss_gpu"], "loss_cpu_oracle_fp32": l_cpu, "rel_err": rel, "tol": 1e-3,
           "ok": rel <= 1e-3, "what": "C2 batch 0, eval mode, same weights and negative draws"}
    if not out["ok"]:
        raise SystemExit(f"bench.py: GPU loss {probe['loss_gpu']} vs CPU oracle {l_cpu}: rel err {rel:.2e} > 1e-3")
    return out


def cpu_retrieval_baseline(D=256, k=200, Bq=4096, Xs=1_000_000):   # the real query batch, 1/10 of the corpus
    from oracle import ref_verbatim
    from oracle.ref_step import torch_topk_baseline
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    items = torch.nn.functional.normalize(torch.randn(Xs, D, generator=g), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(Bq, D, generator=g), dim=-1).to(torch.bfloat16)
    if ref_verbatim.available():
        fn, kind = ref_verbatim.reference_topk, "reference"
        how = "MIPSBruteForceTopK.forward of the staged reference (fp32 mm + torch.topk + id gather, 256 queries per call)"
    else:
        fn, kind = torch_topk_baseline, "port"
        how = "fp32 mm + torch.topk, top_k.py:62-69 restated"
    fn(q[:8], items[:50_000], k)
    t0 = time.perf_counter()
    fn(q, items, k)
    dt = time.perf_counter() - t0
    scale = 10_000_000 / Xs
    return {"value": Bq / (dt * scale), "unit": "queries/s", "cores": torch.get_num_threads(),
            "kind": kind,
            "sample": f"{Bq} queries x {Xs} items ({how}) in {dt:.2f} s, scaled linearly x{scale:.0f} "
                      "to the 10 M corpus"}


def run_reference_arm(args, world, rank):
    if rank != 0:
        return None
    cfg = c2_config(bf16=False)
    ids = synthetic_item_ids(26_744, cfg.num_items)
    sample = PER_GPU_BATCH
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    base = cpu_train_baseline(cfg, ids, sample, steps, warm)
    retr = cpu_retrieval_baseline()
    return {
        "impl": "reference", "metric": "hstu_train_sequences_per_s", "value": base["value"],
        "unit": "sequences/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, world),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "retrieval": {"metric": "topk_retrieval_queries_per_s", "value": retr["value"],
                      "unit": "queries/s", "cpu_baseline": retr,
                      "e2e": {"value": retr["value"], "unit": "queries/s",
                              "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }


def workload_config(cfg, world):
    return {
        "workload": "C2 ml-20m-shaped HSTU training step (BASELINE.json configs[1]): 26,744 items "
                    "(ids <= 131,262), max_seq_len 200 (N=211), 4 HSTU layers, D=256, 4 heads x 64, "
                    "sampled softmax with in-batch negatives R=128 T=0.05, dropout 0.2, AdamW; "
                    "fwd+bwd+optimizer per step; lengths U[20,200]",
        "global_batch": PER_GPU_BATCH * world, "seq_len": cfg.N,
        "parallelism": f"dp{world}" + ("" if world == 1 else (" (DistributedDataParallel)" if os.environ.get("GRB_DDP") == "1" else " (gradients reduced over peer memory: sparse table rows + two-shot all-reduce kernel)")),
        "l2": "per-step working set (134 MB fp32 embedding table + grads + AdamW state, 8 rotating "
              "batches) exceeds the 126 MB L2; no explicit flush",
    }


def _emit(line: dict, out_fd: int) -> None:
    os.write(out_fd, (json.dumps(line) + "\n").encode())


def main():
    # stdout carries exactly ONE JSON line: anything libraries print on fd 1 (e.g. the
    # "NCCL version ..." banner) is diverted to stderr for the duration of the run
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-retrieval", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-long-sequence", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        line = run_reference_arm(args, world, rank)
        if line is not None:
            _emit(line, out_fd)
        return

    world, rank, local = dist_setup(args.gpus)
    train, cfg, ids, model = run_training(args, world, rank, local)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()} if rank == 0 else None
    del model
    torch.cuda.empty_cache()
    retrieval = None if args.skip_retrieval else run_retrieval(args, world, rank, local)
    long_seq = None if args.skip_long_sequence else run_long_sequence(local, world, rank)
    probe = train.pop("_probe", None)
    line = {
        "metric": "hstu_train_sequences_per_s", "value": train["value"], "unit": "sequences/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": train["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(cfg, world),
        "cuda_graphs": os.environ.get("GRB_NO_GRAPHS") != "1",
        "clocks": train["clocks"], "e2e": train["e2e"], "gpu_launches": train["gpu_launches"],
        "roofline": train.get("roofline"), "roofline_others": train.get("roofline_others"),
        "kernels": train["kernels"], "kernels_timed_in": train.get("kernels_timed_in"),
        "final_loss": train["final_loss"], "dropin_eager": train.get("dropin_eager"),
        "long_sequence": long_seq, "retrieval": retrieval,
    }
    if rank == 0:
        if world == 1 and not args.skip_cpu_baseline:
            fp32 = RetrievalConfig(**{**cfg.__dict__, "compute_dtype": None})
            if probe is not None:
                line["loss_check"] = cpu_loss_check(cfg, ids, state, probe)
            line["cpu_baseline"] = cpu_train_baseline(fp32, ids, PER_GPU_BATCH, 3, 1, state)   # ~20-30 s of CPU work
            if retrieval is not None:
                retrieval["cpu_baseline"] = cpu_retrieval_baseline()
        _emit(line, out_fd)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
