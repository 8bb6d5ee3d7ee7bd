#!/usr/bin/env python
"""Stage-level kernel benchmark: the W-forward pruned DFT, fp32 CUDA-core kernel vs TF32 tcgen05 kernel.

Times with CUDA events on the launch stream, an L2 flush (256 MB write) between iterations, and prints
achieved GB/s (algorithmic bytes: read x once, write the kept spectrum once) and TFLOP/s (2*rows*wp*2*m2).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blindno_b200 import ops  # noqa: E402


def bench(rows, wp, m2, prec, iters=20, act=False):
    x = torch.randn(rows, wp, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for _ in range(3):
        ops.stage_wfwd(x, m2, hp=wp, m1=min(m2, wp // 2), prec=prec, act=act)
    times = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.stage_wfwd(x, m2, hp=wp, m1=min(m2, wp // 2), prec=prec, act=act)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    times.sort()
    us = times[len(times) // 2]
    nbytes = rows * (4 * wp + 8 * m2)
    flops = 2.0 * rows * wp * 2 * m2
    return {"rows": rows, "wp": wp, "m2": m2, "act": bool(act),
            "prec": {0: "fp32_ffma", 1: "tf32_tcgen05", 2: "3xtf32_tcgen05"}[prec], "us": us,
            "GBps": nbytes / us / 1e3, "TFLOPs": flops / us / 1e6}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shapes", default="default")
    a = ap.parse_args()
    shapes = [(300 * 4 * 76, 76, 12), (1600 * 4 * 76, 76, 12), (300 * 4 * 100, 100, 12)]
    if a.shapes == "one":
        shapes = shapes[:1]
    for rows, wp, m2 in shapes:
        for act in (False, True):
            for prec in (0, 1, 2):
                print(json.dumps(bench(rows, wp, m2, prec, a.iters, act=act)), flush=True)
