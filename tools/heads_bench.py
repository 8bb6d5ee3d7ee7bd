#!/usr/bin/env python
"""Time ONE output head of the 2D-FPE model (FNO2d: 4 images, width 12, 32 modes, 3 layers, 61 x 61) forward + backward
as a CUDA-graph replay: the few-image, latency-bound half of the train step in isolation (two of these run side by
side in the step).  Prints one JSON line; environment knobs of the library (BDN_PDL, ...) apply."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blindno_b200 import ops  # noqa: E402
from blindno_b200.surface import fno  # noqa: E402

images = int(os.environ.get("HEADS_IMAGES", "4"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.manual_seed(0)
head = fno.FNO2d(modes=32, width=12, n_layers=3, input_dim=12, output_dim=1).cuda().train()
g = torch.Generator().manual_seed(0)
x = torch.randn(images, 61, 61, 12, generator=g).cuda().requires_grad_(True)
gy = torch.randn(images, 61, 61, 1, generator=g).cuda()


def step():
    head.zero_grad(set_to_none=True)
    x.grad = None
    y = head(x)
    y.backward(gy)
    return y


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
n0 = ops.kernel_launches()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
launches = ops.kernel_launches() - n0
for _ in range(10):
    graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    graph.replay()
e1.record()
torch.cuda.synchronize()
print(json.dumps({"images": images, "us_per_fwd_bwd": e0.elapsed_time(e1) * 1e3 / reps, "launches": launches,
                  "pdl": os.environ.get("BDN_PDL", "0"), "tag": os.environ.get("HEADS_TAG", "")}))
