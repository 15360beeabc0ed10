#!/usr/bin/env python
"""Cut ONE graph-replayed train step out of an ncu launch list of bench.py and aggregate it by kernel.

    python tools/step_breakdown.py profiles/r2_launches_bench_train.csv > profiles/r2_step_breakdown.txt
A step = the launches between two consecutive launches of the sampled-softmax forward kernel; the list also
holds eager passes (attribution, drop-in), whose steps are longer: the window with the FEWEST launches is a
graph replay.  Per-launch times under ncu are serialised and cold-cache: compare SHARES."""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    L = [(r[kn], float(r[mv].replace(",", "")) / 1e3) for r in rows[hdr + 1:] if len(r) > mv]
    marks = [i for i, (n, _) in enumerate(L) if "ssl_fwd" in n]
    a, b = min(zip(marks, marks[1:]), key=lambda ab: ab[1] - ab[0])
    step = L[a:b]
    cnt, tot = collections.Counter(), collections.Counter()
    ours = 0.0
    n_ours = 0
    for n, v in step:
        raw = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", raw)).strip()[:70]
        name = name.replace("at::native::", "native::")
        cnt[name] += 1
        tot[name] += v
        if name.startswith("grb::"):
            ours += v
            n_ours += 1
    total = sum(tot.values())
    print(f"# one graph-replayed C2 train step (fwd graph + bwd graph + AdamW) cut out of {sys.argv[1]}")
    print(f"# (launches {a}..{b} of the list: the shortest window between two consecutive ssl_fwd launches)")
    print(f"# {len(step)} launches ({n_ours} of this package), {total:.1f} us of kernel time under ncu (serialised, cold "
          f"cache: compare SHARES); kernels of this package: {100 * ours / total:.1f} %, ATen / CUB: {100 - 100 * ours / total:.1f} %")
    print(f"{'us':>10} {'share':>7} {'launches':>9}  kernel")
    for k, v in tot.most_common(60):
        print(f"{v:10.1f} {100 * v / total:6.1f}% {cnt[k]:9d}  {k}")


if __name__ == "__main__":
    main()
