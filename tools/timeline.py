#!/usr/bin/env python
"""Kernel timeline of graph-replayed train steps (CUPTI through torch.profiler): per-kernel warm durations,
per-stream busy time and the idle gaps between consecutive kernels.  Complements the ncu launch list, whose
durations are cold-cache and serialised."""
import argparse
import collections
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from blindno_b200.parallel import FlatTrainer  # noqa: E402
from blindno_b200.surface import nio  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="2d_FPE")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--bag", type=int, default=75, help="fixed bag size for every profiled step")
    ap.add_argument("--out", default="gpurun_out/timeline.json")
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    model = nio.make_models(wl["variant"])[wl["cls"]](*wl["args"]).to(dev).train()
    trainer = FlatTrainer(model, lr=wl["lr"]).enable_graphs(True)
    grid = bench.make_grid(wl).to(dev)
    x, y = [t.to(dev) for t in bench.make_batches(wl, 1, wl["batch"], seed=0)[0]]
    trainer._draw = lambda n: np.random.choice(n, a.bag)      # same bag size every step: one graph
    for _ in range(5):
        trainer.step(x, grid, y)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            trainer.step(x, grid, y)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    rows = [{"name": e.name[:90], "start_us": e.time_range.start, "dur_us": e.time_range.end - e.time_range.start,
             "stream": getattr(e, "device_index", 0)} for e in evs]
    t0 = rows[0]["start_us"]
    span = rows[-1]["start_us"] + rows[-1]["dur_us"] - t0
    agg = collections.OrderedDict()
    for r in rows:
        k = r["name"].split("(")[0][:60]
        v = agg.setdefault(k, [0, 0.0])
        v[0] += 1
        v[1] += r["dur_us"]
    # union of busy intervals
    busy, cur_s, cur_e = 0.0, None, None
    for r in rows:
        s, e = r["start_us"], r["start_us"] + r["dur_us"]
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    out = {"steps": a.steps, "bag": a.bag, "span_us_per_step": span / a.steps, "gpu_busy_us_per_step": busy / a.steps,
           "idle_us_per_step": (span - busy) / a.steps, "kernels_per_step": len(rows) / a.steps,
           "sum_kernel_us_per_step": sum(r["dur_us"] for r in rows) / a.steps,
           "by_kernel_us_per_step": {k: [v[0] / a.steps, v[1] / a.steps] for k, v in
                                      sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "by_kernel_us_per_step"}))
    for k, v in list(out["by_kernel_us_per_step"].items())[:45]:
        print(f"{k:62s} n={v[0]:5.1f} {v[1]:8.1f} us/step")


if __name__ == "__main__":
    main()
