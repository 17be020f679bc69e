#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel function (share of the captured window)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        k = re.sub(r"\(.*\)$", "", k)
        k = re.sub(r"^void ", "", k)
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms total device time (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:110]}` | {n} | {us:.1f} | {100 * us / tot:.1f}% | {us / n:.2f} |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
