#!/usr/bin/env python
"""Minimal driver for an ncu capture of ONE output head of the 2D-FPE model (FNO2d, 4 images, width 12, 32 modes, 3 layers,
61 x 61 -> padded 76 x 76): forward + backward, eager, fp32 FFMA kernels.  Used with `ncu --set full --import-source on`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blindno_b200.surface import fno  # noqa: E402

torch.manual_seed(0)
head = fno.FNO2d(modes=32, width=12, n_layers=3, input_dim=12, output_dim=1).cuda().train()
g = torch.Generator().manual_seed(0)
x = torch.randn(4, 61, 61, 12, generator=g).cuda().requires_grad_(True)
gy = torch.randn(4, 61, 61, 1, generator=g).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    head.zero_grad(set_to_none=True)
    y = head(x)
    y.backward(gy)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
