"""Autograd-aware Python entry points over the C ABI (include/blindno_b200.h).

torch is used for device memory, streams and autograd bookkeeping only; every FLOP of these
ops runs in libblindno_b200.so.  CPU tensors are rejected -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (FnoGrads, FnoParams, FnoShape, LiftInput, MAX_LAYERS, PREC_FP32, PREC_TF32, PREC_TF32X3, SpectralShape,
                   check, pad_amount)

__all__ = ["FnoSpec", "set_precision", "stage_wfwd", "fno_apply", "spectral_conv", "bag_pool_lift", "adam_step_flat", "kernel_launches",
           "PREC_FP32", "PREC_TF32", "PREC_TF32X3"]


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("blindno_b200 ops run on CUDA tensors only (there is no CPU fallback by design); "
                               f"got a tensor on {t.device}")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 (or complex64 viewed as fp32 pairs), contiguous."""
    if t.is_complex():
        if t.dtype != torch.complex64:
            raise RuntimeError(f"complex weights must be complex64, got {t.dtype}")
        t = torch.view_as_real(t.contiguous())
    if t.dtype != torch.float32:
        raise RuntimeError(f"blindno_b200 ops are fp32, got {t.dtype}")
    return t.contiguous()


def set_precision(module, prec: int):
    """Select the arithmetic of the DFT GEMMs for every FNO net under ``module``: PREC_FP32 (CUDA-core FFMA,
    the 1e-5 parity mode, default) or PREC_TF32 (tcgen05 tensor cores where a stage has a tensor-core
    kernel, TF32 operands / fp32 accumulation; bound 2e-3, tests/test_gpu_parity.py)."""
    if prec not in (PREC_FP32, PREC_TF32, PREC_TF32X3):
        raise ValueError(f"unknown precision {prec}")
    for m in module.modules():
        if hasattr(m, "_spec") and hasattr(m, "spectral_list"):
            m.precision = prec
    return module


def kernel_launches() -> int:
    return int(_lib.lib().bdn_kernel_launches())


def slot_layout(sizes):
    """Offsets of 16-byte aligned slots for tensors of ``sizes`` fp32 elements inside one flat buffer
    (complex views need even offsets; vector loads like 16 bytes).  Returns (offsets, total)."""
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 3) & ~3
    return offs, total


def profile_begin():
    check(_lib.lib().bdn_profile_begin(), "bdn_profile_begin")


def profile_end() -> dict:
    """{"kernel/tag": {"launches": n, "ms": total device ms}} since profile_begin()."""
    import json
    buf = C.create_string_buffer(1 << 20)
    _lib.lib().bdn_profile_end(buf, len(buf))
    return json.loads(buf.value.decode() or "{}")


def stage_wfwd(x: torch.Tensor, m2: int, *, hp: int = 1, m1: int = 0, act: bool = False, prec: int = PREC_FP32):
    """One stage on its own (tests, kernel benchmarks): pruned forward DFT along the last axis of
    x [rows, wp] -> complex64 [rows, m2].  ``prec=PREC_TF32`` runs the tcgen05 tensor-core kernel."""
    _need_cuda(x)
    xc = _f32c(x)
    rows, wp = xc.shape
    out = torch.empty(rows, m2, 2, dtype=torch.float32, device=xc.device)
    with torch.cuda.device(xc.device):
        check(_lib.lib().bdn_stage_wfwd(hp, wp, m1, m2, rows, _ptr(xc), _ptr(out), int(act), prec, _stream()),
              "bdn_stage_wfwd")
    return torch.view_as_complex(out)


# ---------------------------------------------------------------------------------------------
# one FNO net
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FnoSpec:
    """Static description of an FNO net (FNO1d/FNO2d ctor arguments of the reference)."""
    ndim: int
    c_in: int
    width: int
    c_out: int
    n_layers: int
    modes1: int        # 2-D kept rows per block (0 in 1-D)
    modes2: int        # kept columns (1-D: modes)
    hidden: int = 128
    prec: int = PREC_FP32

    @property
    def n_params(self) -> int:
        return 6 + (4 if self.ndim == 2 else 3) * self.n_layers

    def split(self, params: Sequence[torch.Tensor]):
        """fc0_w, fc0_b, conv_w[n], conv_b[n], spec_w1[n], (spec_w2[n]), fc1_w, fc1_b, fc2_w, fc2_b"""
        n = self.n_layers
        it = iter(params)
        fc0_w, fc0_b = next(it), next(it)
        conv_w = [next(it) for _ in range(n)]
        conv_b = [next(it) for _ in range(n)]
        w1 = [next(it) for _ in range(n)]
        w2 = [next(it) for _ in range(n)] if self.ndim == 2 else [None] * n
        fc1_w, fc1_b, fc2_w, fc2_b = next(it), next(it), next(it), next(it)
        return fc0_w, fc0_b, conv_w, conv_b, w1, w2, fc1_w, fc1_b, fc2_w, fc2_b


def _fill_params(spec: FnoSpec, tensors) -> FnoParams:
    fc0_w, fc0_b, conv_w, conv_b, w1, w2, fc1_w, fc1_b, fc2_w, fc2_b = spec.split(tensors)
    p = FnoParams()
    p.fc0_w, p.fc0_b = _ptr(fc0_w), _ptr(fc0_b)
    for k in range(spec.n_layers):
        p.conv_w[k], p.conv_b[k] = _ptr(conv_w[k]), _ptr(conv_b[k])
        p.spec_w1[k], p.spec_w2[k] = _ptr(w1[k]), _ptr(w2[k])
    p.fc1_w, p.fc1_b, p.fc2_w, p.fc2_b = _ptr(fc1_w), _ptr(fc1_b), _ptr(fc2_w), _ptr(fc2_b)
    return p


def _make_shape(spec: FnoSpec, images: int, h: int, w: int) -> FnoShape:
    pad_h = pad_amount(h) if spec.ndim == 2 else 0
    pad_w = pad_amount(w)
    s = FnoShape()
    s.ndim, s.images, s.c_in, s.width, s.c_out, s.hidden = spec.ndim, images, spec.c_in, spec.width, spec.c_out, spec.hidden
    s.n_layers, s.h, s.w, s.hp, s.wp = spec.n_layers, h, w, h + pad_h, w + pad_w
    # Q4: the reference crops H by the W-derived pad and W by the H-derived pad (2-D only)
    s.out_h = s.hp - pad_w if spec.ndim == 2 else 1
    s.out_w = s.wp - pad_h if spec.ndim == 2 else w
    s.m1, s.m2, s.prec = spec.modes1, spec.modes2, spec.prec
    if s.out_h < 1 or s.out_w < 1:
        raise RuntimeError(f"crop leaves an empty grid ({s.out_h} x {s.out_w})")
    return s


class _FnoFn(torch.autograd.Function):
    """FNO1d/FNO2d.forward (+ optionally the bag mean and detached lift on its output)."""

    @staticmethod
    def forward(ctx, spec: FnoSpec, x_cl, bags, idx, grid, pool_w0, pool_b0, sink, *params):
        L = _lib.lib()
        _need_cuda(x_cl, bags, grid, *params)
        flat = [_f32c(p.detach()) for p in params]
        if len(flat) != spec.n_params:
            raise RuntimeError(f"expected {spec.n_params} parameter tensors, got {len(flat)}")
        lift = LiftInput()
        if x_cl is not None:
            x_cl = _f32c(x_cl.detach())
            if spec.ndim == 2:
                images, h, w, cin = x_cl.shape
            else:
                (images, w, cin), h = x_cl.shape, 1
            if cin != spec.c_in:
                raise RuntimeError(f"input has {cin} features, fc0 expects {spec.c_in}")
            lift.x_cl = _ptr(x_cl)
            n_bags = n_keep = 0
            dev = x_cl.device
        else:
            bags, grid = _f32c(bags.detach()), _f32c(grid.detach())
            if spec.ndim == 2:
                n_bags, bag_len, h, w = bags.shape
            else:
                (n_bags, bag_len, w), h = bags.shape, 1
            if idx is not None:
                idx = idx.to(device=bags.device, dtype=torch.int32).contiguous()
                n_keep = idx.numel()
            else:
                n_keep = bag_len
            gd = grid.shape[-1]
            if grid.numel() != h * w * gd or 1 + gd != spec.c_in:
                raise RuntimeError(f"grid shape {tuple(grid.shape)} does not match bags {tuple(bags.shape)}")
            images = n_bags * n_keep
            lift.bags, lift.idx, lift.grid = _ptr(bags), _ptr(idx), _ptr(grid)
            lift.n_bags, lift.bag_len, lift.n_keep, lift.grid_dim = n_bags, bag_len, n_keep, gd
            dev = bags.device
        shape = _make_shape(spec, images, h, w)
        pooled = pool_w0 is not None
        if pooled:
            if x_cl is not None or spec.c_out != 1:
                raise RuntimeError("bag pooling needs the bag input form and a scalar FNO output")
            pool_w0, pool_b0 = _f32c(pool_w0.detach()), _f32c(pool_b0.detach())

        need_grad = any(ctx.needs_input_grad)
        with torch.cuda.device(dev):
            ws_bytes = L.bdn_fno_workspace_bytes(C.byref(shape))
            if ws_bytes == 0 and images > 0:
                check(-1, "bdn_fno_workspace_bytes")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            z_saved = xs_saved = None
            if need_grad:
                z_saved = torch.empty(L.bdn_fno_act_floats(C.byref(shape)), dtype=torch.float32, device=dev)
                xs_saved = torch.empty(L.bdn_fno_spec_floats(C.byref(shape)), dtype=torch.float32, device=dev)
            out_shape = (images, shape.out_h, shape.out_w, spec.c_out) if spec.ndim == 2 else (images, shape.out_w, spec.c_out)
            out = torch.empty(out_shape, dtype=torch.float32, device=dev)
            cparams = _fill_params(spec, flat)
            check(L.bdn_fno_forward(C.byref(shape), C.byref(cparams), C.byref(lift), _ptr(out), _ptr(z_saved),
                                    _ptr(xs_saved), _ptr(ws), ws_bytes, _stream()), "bdn_fno_forward")
            if pooled:
                npix = shape.out_h * shape.out_w
                if npix != h * w:
                    raise RuntimeError("bag pooling needs the FNO output on the input grid (square 2-D grids)")
                width0 = pool_w0.shape[0]
                lifted_shape = (n_bags, h, w, width0) if spec.ndim == 2 else (n_bags, w, width0)
                lifted = torch.empty(lifted_shape, dtype=torch.float32, device=dev)
                check(L.bdn_bag_pool_lift_forward(_ptr(out), _ptr(grid), _ptr(pool_w0), _ptr(pool_b0), _ptr(lifted),
                                                  n_bags, n_keep, npix, gd, width0, _stream()),
                      "bdn_bag_pool_lift_forward")
                result = lifted
            else:
                result = out
        ctx.spec, ctx.shape, ctx.lift = spec, shape, lift
        ctx.keep = (x_cl, bags, idx, grid, pool_w0, flat, z_saved, xs_saved)   # keeps the raw pointers alive
        ctx.pooled, ctx.n_bags, ctx.n_keep = pooled, n_bags, n_keep
        ctx.sink = sink
        ctx.param_meta = [(p.shape, p.is_complex()) for p in params]
        return result

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        spec, shape, lift = ctx.spec, ctx.shape, ctx.lift
        x_cl, bags, idx, grid, pool_w0, flat, z_saved, xs_saved = ctx.keep
        dev = z_saved.device
        g = _f32c(g)
        with torch.cuda.device(dev):
            if ctx.pooled:
                npix = shape.out_h * shape.out_w
                gpool = torch.empty(ctx.n_bags * npix, dtype=torch.float32, device=dev)
                check(L.bdn_bag_pool_lift_backward(_ptr(g), _ptr(pool_w0), _ptr(gpool), ctx.n_bags, npix,
                                                   lift.grid_dim, pool_w0.shape[0], _stream()),
                      "bdn_bag_pool_lift_backward")
                g = gpool
            sizes = [t.numel() for t in flat]
            offs, total = slot_layout(sizes)
            if ctx.sink is not None:
                # the trainer's flat gradient buffer: kernels accumulate into it, autograd sees no grads
                if ctx.sink.numel() != total or ctx.sink.dtype != torch.float32 or not ctx.sink.is_contiguous():
                    raise RuntimeError("gradient sink does not match this net's slot layout")
                gflat = ctx.sink
            else:
                gflat = torch.zeros(total, dtype=torch.float32, device=dev)
            gviews = [gflat[o:o + n] for o, n in zip(offs, sizes)]
            cgrads = _fill_params(spec, gviews)
            gx = None
            if x_cl is not None and ctx.needs_input_grad[1]:
                gx = torch.empty_like(x_cl)
            ws_bytes = L.bdn_fno_workspace_bytes(C.byref(shape))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            cparams = _fill_params(spec, flat)
            check(L.bdn_fno_backward(C.byref(shape), C.byref(cparams), C.byref(lift), _ptr(g), int(ctx.pooled),
                                     max(ctx.n_keep, 1), _ptr(z_saved), _ptr(xs_saved), C.byref(cgrads), _ptr(gx),
                                     _ptr(ws), ws_bytes, _stream()), "bdn_fno_backward")
        grads = []
        for view, (shp, is_c), need in zip(gviews, ctx.param_meta, ctx.needs_input_grad[8:]):
            if not need or ctx.sink is not None:
                grads.append(None)
            elif is_c:
                grads.append(torch.view_as_complex(view.view(*shp, 2)))
            else:
                grads.append(view.view(shp))
        return (None, gx, None, None, None, None, None, None, *grads)


def fno_apply(spec: FnoSpec, params: Sequence[torch.Tensor], *, x_cl=None, bags=None, idx=None, grid=None,
              pool=None, grad_sink=None) -> torch.Tensor:
    """Run one FNO net.

    ``x_cl``: channels-last input [images, (h,) w, c_in]; or ``bags`` [B, L0, (h,) w] + ``grid`` [(h,) w, d]
    (+ optional int ``idx`` of kept snapshots) for the per-snapshot NIO-FNO encoder, whose input
    concat(snapshot, grid) is never materialised.  ``pool=(fc0.weight, fc0.bias)`` additionally applies the
    bag mean and the detached lift, returning [B, (h,) w, width] instead of the per-snapshot outputs.
    ``grad_sink``: a flat fp32 buffer laid out by ``slot_layout`` over ``params``; when given, the backward
    kernels accumulate the parameter gradients straight into it (the data-parallel trainer all-reduces that
    buffer in one call) and autograd receives no parameter gradients.
    """
    pw, pb = pool if pool is not None else (None, None)
    return _FnoFn.apply(spec, x_cl, bags, idx, grid, pw, pb, grad_sink, *params)


# ---------------------------------------------------------------------------------------------
# one spectral convolution
# ---------------------------------------------------------------------------------------------
class _SpectralFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, w2, ndim):
        L = _lib.lib()
        _need_cuda(x, w1, w2)
        xc, w1c = _f32c(x.detach()), _f32c(w1.detach())
        w2c = _f32c(w2.detach()) if w2 is not None else None
        s = SpectralShape()
        s.ndim, s.prec = ndim, PREC_FP32
        if ndim == 2:
            s.images, s.c_in, s.hp, s.wp = xc.shape
            ci, s.c_out, s.m1, s.m2 = w1c.shape[:4]
        else:
            (s.images, s.c_in, s.wp), s.hp, s.m1 = xc.shape, 1, 0
            ci, s.c_out, s.m2 = w1c.shape[:3]
        if ci != s.c_in:
            raise RuntimeError(f"weights expect {ci} input channels, x has {s.c_in}")
        dev = xc.device
        with torch.cuda.device(dev):
            ws_bytes = L.bdn_spectral_workspace_bytes(C.byref(s))
            if ws_bytes == 0:
                check(-1, "bdn_spectral_workspace_bytes")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            K = 2 * s.m1 if ndim == 2 else 1
            xs = torch.empty(s.images * s.c_in * K * s.m2 * 2, dtype=torch.float32, device=dev) \
                if any(ctx.needs_input_grad) else None
            y = torch.empty((s.images, s.c_out) + tuple(xc.shape[2:]), dtype=torch.float32, device=dev)
            check(L.bdn_spectral_forward(C.byref(s), _ptr(xc), _ptr(w1c), _ptr(w2c), _ptr(y), _ptr(xs), _ptr(ws),
                                         ws_bytes, _stream()), "bdn_spectral_forward")
        ctx.s, ctx.keep = s, (w1c, w2c, xs)
        ctx.meta = (w1.shape, w1.is_complex())
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.lib()
        s = ctx.s
        w1c, w2c, xs = ctx.keep
        gy = _f32c(gy)
        dev = gy.device
        with torch.cuda.device(dev):
            ws_bytes = L.bdn_spectral_workspace_bytes(C.byref(s))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            gx = torch.empty((s.images, s.c_in) + tuple(gy.shape[2:]), dtype=torch.float32, device=dev) \
                if ctx.needs_input_grad[0] else None
            gw1 = torch.empty_like(w1c)
            gw2 = torch.empty_like(w2c) if w2c is not None else None
            check(L.bdn_spectral_backward(C.byref(s), _ptr(gy), _ptr(xs), _ptr(w1c), _ptr(w2c), _ptr(gx), _ptr(gw1),
                                          _ptr(gw2), _ptr(ws), ws_bytes, _stream()), "bdn_spectral_backward")
        shp, is_c = ctx.meta
        if is_c:
            gw1 = torch.view_as_complex(gw1)
            gw2 = torch.view_as_complex(gw2) if gw2 is not None else None
        return gx, gw1, gw2, None


def spectral_conv(x: torch.Tensor, w1: torch.Tensor, w2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SpectralConv2d.forward (x [B,C,H,W], w1/w2 real pairs [...,2] or complex64) or, with ``w2=None``,
    SpectralConv1d.forward (x [B,C,N], w1 complex64 [Ci,Co,m]; DC bin halved as in the reference)."""
    return _SpectralFn.apply(x, w1, w2, 2 if w2 is not None else 1)


# ---------------------------------------------------------------------------------------------
# bag mean + detached lift on its own (the NIO models feed it DeepONet outputs)
# ---------------------------------------------------------------------------------------------
class _PoolLiftFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, grid, w0, b0):
        L = _lib.lib()
        _need_cuda(s, grid, w0, b0)
        sc, gc, w0c, b0c = _f32c(s.detach()), _f32c(grid.detach()), _f32c(w0.detach()), _f32c(b0.detach())
        n_bags, n_keep = sc.shape[:2]
        gshape = tuple(sc.shape[2:])
        npix = 1
        for d in gshape:
            npix *= d
        gd, width = gc.shape[-1], w0c.shape[0]
        out = torch.empty((n_bags,) + gshape + (width,), dtype=torch.float32, device=sc.device)
        with torch.cuda.device(sc.device):
            check(L.bdn_bag_pool_lift_forward(_ptr(sc), _ptr(gc), _ptr(w0c), _ptr(b0c), _ptr(out), n_bags, n_keep,
                                              npix, gd, width, _stream()), "bdn_bag_pool_lift_forward")
        ctx.keep, ctx.dims = (w0c,), (n_bags, n_keep, npix, gd, width, tuple(sc.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        (w0c,) = ctx.keep
        n_bags, n_keep, npix, gd, width, sshape = ctx.dims
        g = _f32c(g)
        gpool = torch.empty(n_bags * npix, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(L.bdn_bag_pool_lift_backward(_ptr(g), _ptr(w0c), _ptr(gpool), n_bags, npix, gd, width, _stream()),
                  "bdn_bag_pool_lift_backward")
        gs = (gpool.view(n_bags, 1, npix) / n_keep).expand(n_bags, n_keep, npix).reshape(sshape)
        return gs, None, None, None


def bag_pool_lift(s, grid, w0, b0):
    """fc0([grid, mean_l s_l]) with fc0 detached: s [B,L,*g], grid [*g,d] -> [B,*g,width]."""
    return _PoolLiftFn.apply(s, grid, w0, b0)


def adam_step_flat(param, grad, exp_avg, exp_avg_sq, *, lr, betas=(0.9, 0.999), eps=1e-8, step, grad_scale=1.0):
    """torch.optim.Adam's update over one flat fp32 buffer, in place, one kernel."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    with torch.cuda.device(param.device):
        check(_lib.lib().bdn_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(),
                                       lr, betas[0], betas[1], eps, step, grad_scale, _stream()), "bdn_adam_step")
