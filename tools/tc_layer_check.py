#!/usr/bin/env python
"""GPU bring-up check of the fused tensor-core layer kernels (csrc/tc_layer.cu): one 2-D layer, forward and
backward, in the TF32 / 3xTF32 modes against the FFMA kernels (themselves pinned to the oracle by tests/) and
against an fp64 evaluation of the same layer with torch.fft on the CPU.  Prints one JSON line per case."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blindno_b200 import _lib, ops  # noqa: E402

CASES = [
    # images, width, hp, wp, m1, m2
    (8, 4, 76, 76, 12, 12),        # few images: one channel plane per CTA
    (300, 4, 76, 76, 12, 12),      # the per-snapshot net: whole images per CTA (cg = 4)
    (4, 12, 76, 76, 32, 32),       # the output heads
    (37, 4, 76, 76, 12, 12),
    (160, 8, 40, 52, 6, 10),       # rectangular, cg = 4 of 8 channels
    (3, 6, 28, 36, 5, 7),          # odd m1: K = 10 padded to 16
    (200, 4, 100, 100, 12, 12),    # 2D-NC per-snapshot net
    (4, 12, 100, 100, 32, 32),     # 2D-NC heads
]


def ref64(z, w1, w2, cw, cb, gelu_in, m1, m2):
    """fp64 layer body exactly as 2d_FPE/FNOModules.py:156-178, 226-232 composes it (pre-activation in / out)."""
    x = torch.nn.functional.gelu(z) if gelu_in else z
    b, c, hp, wp = x.shape
    xf = torch.fft.rfft2(x)
    out = torch.zeros(b, c, hp, wp // 2 + 1, dtype=torch.complex128)
    out[:, :, :m1, :m2] = torch.einsum("bixy,ioxy->boxy", xf[:, :, :m1, :m2], torch.view_as_complex(w1))
    out[:, :, -m1:, :m2] = torch.einsum("bixy,ioxy->boxy", xf[:, :, -m1:, :m2], torch.view_as_complex(w2))
    y = torch.fft.irfft2(out, s=(hp, wp))
    return y + torch.einsum("oi,bihw->bohw", cw, x) + cb[None, :, None, None]


def run(case, prec, seed=0):
    images, c, hp, wp, m1, m2 = case
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(images, c, hp, wp, generator=g)
    scale = 1.0 / (c * c)
    w1 = scale * torch.rand(c, c, m1, m2, 2, generator=g)
    w2 = scale * torch.rand(c, c, m1, m2, 2, generator=g)
    cw = torch.randn(c, c, generator=g) / c ** 0.5
    cb = torch.randn(c, generator=g)
    gy = torch.randn(images, c, hp, wp, generator=g)
    res = {}
    for gelu_in in (False, True):
        leaves = [t.cuda().requires_grad_(True) for t in (z, w1, w2, cw, cb)]
        out = ops.fno_layer(leaves[0], leaves[1], leaves[2], leaves[3].view(c, c, 1, 1), leaves[4], gelu_in, prec)
        out.backward(gy.cuda())
        torch.cuda.synchronize()
        res[gelu_in] = [out.detach().cpu()] + [t.grad.detach().cpu() for t in leaves]
    return res


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="")
    ap.add_argument("--ref64", action="store_true", help="also compare with an fp64 torch.fft evaluation (small cases)")
    a = ap.parse_args()
    sel = [int(i) for i in a.cases.split(",")] if a.cases else range(len(CASES))
    names = ["z_out", "gz_in", "g_w1", "g_w2", "g_conv_w", "g_conv_b"]
    for i in sel:
        case = CASES[i]
        images, c, hp, wp, m1, m2 = case
        s = ops._stage_shape(2, images, c_in=c, width=c, h=hp, w=wp, hp=hp, wp=wp, out_h=hp, out_w=wp, m1=m1, m2=m2, prec=2)
        import ctypes
        path = _lib.lib().bdn_fno_layer_path(ctypes.byref(s))
        base = run(case, 0)
        line = {"case": case, "tc_path": int(path), "swap": os.environ.get("BDN_TC_DESC_SWAP", "0")}
        for prec, tag in ((2, "x3"), (1, "tf32")):
            ops.profile_begin()
            got = run(case, prec)
            tags = sorted(ops.profile_end().keys())
            line[f"kernels_{tag}"] = [t for t in tags if t.startswith(("tc_", "wfwd", "core2d", "winv"))]
            for gelu_in in (False, True):
                line[f"{tag}_gelu{int(gelu_in)}"] = {n: float("%.3g" % rel(g_, b_)) for n, g_, b_ in zip(names, got[gelu_in], base[gelu_in])}
        if a.ref64 and images * c * hp * wp <= 4_000_000:
            g = torch.Generator().manual_seed(0)
            z = torch.randn(images, c, hp, wp, generator=g)
            scale = 1.0 / (c * c)
            w1 = scale * torch.rand(c, c, m1, m2, 2, generator=g)
            w2 = scale * torch.rand(c, c, m1, m2, 2, generator=g)
            cw = torch.randn(c, c, generator=g) / c ** 0.5
            cb = torch.randn(c, generator=g)
            gy = torch.randn(images, c, hp, wp, generator=g)
            leaves = [t.double().requires_grad_(True) for t in (z, w1, w2, cw, cb)]
            want = ref64(*leaves, True, m1, m2)
            want.backward(gy.double())
            want = [want.detach()] + [t.grad for t in leaves]
            got = run(case, 2)[True]
            line["x3_vs_fp64_gelu1"] = {n: float("%.3g" % rel(g_, w_)) for n, g_, w_ in zip(names, got, want)}
            line["ffma_vs_fp64_gelu1"] = {n: float("%.3g" % rel(g_, w_)) for n, g_, w_ in zip(names, base[True], want)}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
