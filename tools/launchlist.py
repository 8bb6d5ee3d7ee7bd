#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid, block)."""
import collections
import csv
import sys


def load(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def main(path, width=78):
    rows = load(path)
    agg = collections.OrderedDict()
    for r in rows:
        k = (r["Kernel Name"][:width], r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':{width}s} {'grid':16s} {'block':12s} {'n':>4s} {'avg us':>9s} {'total us':>10s} share")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[0]:{width}s} {k[1]:16s} {k[2]:12s} {a[0]:4d} {a[1] / a[0]:9.1f} {a[1]:10.1f} {100 * a[1] / tot:5.1f}%")
    print(f"total {tot:.1f} us over {len(rows)} launches")


if __name__ == "__main__":
    main(sys.argv[1])
