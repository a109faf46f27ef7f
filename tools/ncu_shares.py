#!/usr/bin/env python
"""Per-kernel totals and shares from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>`).

    python tools/ncu_shares.py gpurun_out/launches.csv [--skip-first N] [--between SUBSTR I J] > profiles/rNN_ncu_launch_shares.csv

--between SUBSTR I J keeps the launches after the I-th and up to (including) the J-th launch whose kernel name contains
SUBSTR (1-based) -- e.g. `--between reduce_levels_kernel 1 2` is exactly the second training step of a bench.py capture
(reduce_levels_kernel is the last kernel of a reverse sweep; the optimizer's few element-wise kernels after it are the
head of the kept range instead of its tail).

Times under ncu are serialised and cold-cache: compare SHARES with bench.py's kernel_breakdown_ms, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)                       # drop the argument list
    name = name.replace("void ", "").replace("dadmm::", "")
    m = re.match(r"([\w:]+)(<[^>]{0,40})?", name)
    return (m.group(1) + (m.group(2) + ">" if m.group(2) else "")) if m else name[:60]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip-first") + 1]) if "--skip-first" in sys.argv else 0
    lines = open(path, newline="").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"][skip:]
    if "--between" in sys.argv:
        i = sys.argv.index("--between")
        sub, a, b = sys.argv[i + 1], int(sys.argv[i + 2]), int(sys.argv[i + 3])
        hits = [j for j, r in enumerate(rows) if sub in r["Kernel Name"]]
        rows = rows[hits[a - 1] + 1:hits[b - 1] + 1]
    agg = OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0.0, 0])
        a[0] += ms
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print("ms,launches,share,kernel")
    for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:.3f},{n},{100 * ms / tot:.1f}%,{k}")
    print(f"# total {tot:.3f} ms over {sum(a[1] for a in agg.values())} launches")


if __name__ == "__main__":
    main()
