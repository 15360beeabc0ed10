#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py profiles/r1_launches_bench_train.csv > profiles/r1_launches_bench_train_summary.txt
Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolute times."""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    cnt, tot = collections.Counter(), collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        raw = r[kn].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", raw)).strip()[:70]
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu].replace("second", "s").replace("n", "n"), 1e-3) if r[mu] in ("ns", "us", "ms") else 1e-3
        cnt[name] += 1
        tot[name] += v
    total = sum(tot.values())
    print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {total / 1e3:.2f} ms of kernel time (serialised, cold cache)")
    print(f"{'us total':>12} {'share':>7} {'launches':>9}  kernel")
    for k, v in tot.most_common(45):
        print(f"{v:12.1f} {100 * v / total:6.1f}% {cnt[k]:9d}  {k}")


if __name__ == "__main__":
    main()
