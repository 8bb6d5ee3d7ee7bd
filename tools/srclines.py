#!/usr/bin/env python
"""Per CUDA source line summary of an `ncu -i rep --page source --csv --print-source cuda,sass` dump: sampled stall
counts and executed warp instructions per source line, top lines first.
usage: srclines.py dump.csv kernel-substring [instance] [top]"""
import csv
import sys


def sections(path):
    out, cur, hdr, fpath = [], None, None, None
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1]
        elif row[0] == "Function Name":
            name = row[1]
            if cur is None or cur["name"] != name or fpath in cur["seen"]:
                cur = {"name": name, "rows": [], "seen": set()}
                out.append(cur)
            cur["seen"].add(fpath)
            hdr = None
        elif row[0] == "Line No":
            hdr = row
        elif cur is not None and hdr is not None and len(row) == len(hdr) and row[0] != "":
            d = {}
            for h, v in zip(hdr, row):      # "Source" appears twice: keep the first (the CUDA line)
                d.setdefault(h, v)
            d["file"] = fpath
            cur["rows"].append(d)
    return out


def main():
    path, pat = sys.argv[1], sys.argv[2]
    inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    ks = [k for k in sections(path) if pat in k["name"]]
    k = ks[inst]
    rows = []
    for r in k["rows"]:
        try:
            s = int(r["# Samples"])
            n = int(r["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        if s or n:
            stalls = {c[6:]: int(r[c]) for c in r if c.startswith("stall_") and "Not Issued" not in c and r[c].isdigit() and int(r[c])}
            rows.append((s, n, r["file"].split("/")[-1], r["Line No"], r["Source"].strip()[:100], stalls))
    tot_s = sum(r[0] for r in rows)
    tot_n = sum(r[1] for r in rows)
    print(k["name"], "instances:", len(ks), "samples", tot_s, "warp-instr", tot_n)
    for s, n, f, ln, src, st in sorted(rows, key=lambda r: -r[0])[:top]:
        top3 = " ".join(f"{a}:{b}" for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100.0 * s / max(tot_s, 1):5.1f}% smp {100.0 * n / max(tot_n, 1):5.1f}% ins  {f}:{ln:>4}  {src}   [{top3}]")


if __name__ == "__main__":
    main()
