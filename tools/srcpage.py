#!/usr/bin/env python
"""Summarise `ncu --page source --csv` (SASS view): per kernel, group instructions by execution count
(instructions of one loop share a trip count) and list opcode mixes and stall samples per group."""
import collections
import csv
import sys


def kernels(path):
    cur, hdr, out = None, None, []
    for row in csv.reader(open(path)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            out.append(cur)
            hdr = None
        elif row and row[0] == "Address":
            hdr = row
        elif cur is not None and hdr is not None and len(row) >= 7:
            cur["rows"].append(dict(zip(hdr, row)))
    return out


def main(path, pattern, which=0, top=14):
    ks = [k for k in kernels(path) if pattern in k["name"]]
    k = ks[which]
    print(k["name"], "instances matching:", len(ks))
    groups = collections.OrderedDict()
    tot = 0
    for r in k["rows"]:
        n = int(r["Instructions Executed"])
        tot += n
        g = groups.setdefault(n, {"count": 0, "ops": collections.Counter(), "samples": 0, "first": r["Address"][-5:]})
        g["count"] += 1
        g["ops"][r["Source"].split()[0 if not r["Source"].strip().startswith("@") else 1].split(".")[0]] += 1
        g["samples"] += int(r["# Samples"])
    print("total warp-instructions", tot)
    for n, g in sorted(groups.items(), key=lambda kv: -kv[0] * kv[1]["count"])[:top]:
        ops = " ".join(f"{o}:{c}" for o, c in g["ops"].most_common(9))
        print(f"  exec/inst {n:9d} x {g['count']:4d} insts = {100.0 * n * g['count'] / tot:5.1f}%  samples {g['samples']:6d}  @{g['first']}  {ops}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
