#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full --csv --page raw --log-file X.csv` capture, grouped by kernel
(name, grid, block) and averaged over its launches: duration, DRAM bytes and throughput, issue-slot and
tensor-pipe utilisation, the dominant stall reasons.  Usage: tools/ncucsv.py X.csv [--json out.json]"""
import collections
import csv
import json
import sys

COLS = {
    "us": ("gpu__time_duration.sum", 1e-3),                     # ns -> us (the raw page reports ns)
    "rdMB": ("dram__bytes_read.sum", None),
    "wrMB": ("dram__bytes_write.sum", None),
    "dram%": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "issue%": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "tensor%": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "fma%": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "xu%": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    "warps%": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "regs": ("launch__registers_per_thread", 1),
    "longSB": ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
    "shortSB": ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1),
    "barrier": ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1),
    "Minst": ("smsp__inst_executed.sum", 1e-6),
}


def load(path):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.reader(lines[start:]))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, [r for r in data if len(r) == len(hdr)]


def main(path, json_out=None):
    hdr, units, data = load(path)
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        if name not in col:
            return float("nan")
        try:
            v = float(r[col[name]].replace(",", ""))
        except ValueError:
            return float("nan")
        u = units[col[name]].lower()
        if name.startswith("dram__bytes"):                      # normalise to MB whatever unit ncu chose
            v *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
        return v

    agg = collections.OrderedDict()
    for r in data:
        key = (r[col["Kernel Name"]].split("(")[0][:44], r[col["Grid Size"]].replace(" ", ""), r[col["Block Size"]].replace(" ", ""))
        a = agg.setdefault(key, {"n": 0, **{k: 0.0 for k in COLS}})
        a["n"] += 1
        for k, (name, scale) in COLS.items():
            v = val(r, name)
            if scale not in (None, 1) and name != "gpu__time_duration.sum":
                v *= scale
            a[k] += v
    names = list(COLS)
    print("%-44s %-12s %-11s %3s " % ("kernel", "grid", "block", "n") + " ".join("%8s" % n for n in names) + "   GB/s")
    out = []
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        avg = {k: a[k] / n for k in names}
        gbs = (avg["rdMB"] + avg["wrMB"]) * 1e6 / (avg["us"] * 1e-6) / 1e9 if avg["us"] > 0 else float("nan")
        print("%-44s %-12s %-11s %3d " % (*key, n) + " ".join("%8.2f" % avg[k] for k in names) + " %7.0f" % gbs)
        out.append({"kernel": key[0], "grid": key[1], "block": key[2], "launches": n, **avg, "dram_GBps": gbs})
    if json_out:
        json.dump(out, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None)
