#!/usr/bin/env python
"""Minimal driver for an ncu capture of the fused tensor-core layer kernels: one 2-D layer, forward + backward,
3xTF32, at the per-snapshot-net shape (300 x C=4) or the heads' shape (4 x C=12)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blindno_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "snap"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 2
SHAPES = {"snap": (300, 4, 76, 76, 12, 12), "snap32": (2400, 4, 76, 76, 12, 12), "heads": (4, 12, 76, 76, 32, 32),
          "heads32": (32, 12, 76, 76, 32, 32)}
images, c, hp, wp, m1, m2 = SHAPES[which]
g = torch.Generator().manual_seed(0)
z = torch.randn(images, c, hp, wp, generator=g).cuda().requires_grad_(True)
w1 = (torch.rand(c, c, m1, m2, 2, generator=g) / (c * c)).cuda().requires_grad_(True)
w2 = (torch.rand(c, c, m1, m2, 2, generator=g) / (c * c)).cuda().requires_grad_(True)
cw = (torch.randn(c, c, 1, 1, generator=g) / c ** 0.5).cuda().requires_grad_(True)
cb = torch.randn(c, generator=g).cuda().requires_grad_(True)
gy = torch.randn(images, c, hp, wp, generator=g).cuda()
for _ in range(2):
    out = ops.fno_layer(z, w1, w2, cw, cb, True, prec)
    out.backward(gy)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
