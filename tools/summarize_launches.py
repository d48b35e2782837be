#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    total = sum(r[1] for r in rows)
    agg = OrderedDict()
    for name, ns, _, _ in rows:
        short = name.split("(")[0]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    print(f"# {path}: {len(rows)} launches, {total/1e3:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'n':>4s} {'us':>10s} {'share':>7s}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:4d} {ns/1e3:10.1f} {ns/total*100:6.1f}%")
    if "-v" in sys.argv:
        for i, (name, ns, g, b) in enumerate(rows):
            print(f"{i:3d} {ns/1e3:9.1f} us  {g:>14s} {b:>12s}  {name.split('(')[0][:60]}")


if __name__ == "__main__":
    main(sys.argv[1])
