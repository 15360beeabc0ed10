#!/usr/bin/env python
"""Summarise an ncu report (one kernel) into a small text file for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_name.txt
Needs `ncu` on PATH (no GPU): reads the raw page for the headline metrics and the source page
for the per-opcode stall-sample histogram."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = [f"# ncu summary of {rep}", ""]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    lines.append(f"kernel: {vals[hdr.index('Kernel Name')]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"{k} = {vals[i]} {units[i]}")
    src = [r for r in page(rep, "source")[2:] if len(r) > 5 and r[4].strip().isdigit() and r[5].strip().isdigit()]
    tot_s = sum(int(r[4]) for r in src) or 1
    tot_i = sum(int(r[5]) for r in src) or 1
    ops = collections.defaultdict(lambda: [0, 0])
    for r in src:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op][0] += int(r[4]); ops[op][1] += int(r[5])
    lines += ["", f"SASS instructions executed (warp-level): {tot_i}; stall samples: {tot_s}",
              "opcode        samples%   executed%"]
    for k, (s, i) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:24]:
        lines.append(f"{k:12s} {100 * s / tot_s:7.1f}   {100 * i / tot_i:7.1f}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
