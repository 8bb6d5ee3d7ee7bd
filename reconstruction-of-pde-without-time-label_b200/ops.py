"""PyTorch custom ops (``torch.ops.blindno_b200.*``) over the C ABI (include/blindno_b200.h).

The reference has no operator registry of its own: its interface for this path is the ``forward`` of a
few nn.Modules.  This module registers one custom op per fused stage with
``torch.library`` -- schema, CUDA implementation (a ctypes call into libblindno_b200.so), a shape-only
fake/meta implementation and an autograd formula -- and the drop-in modules in ``surface/`` call those
ops.  torch provides device memory, the current stream, the dispatcher and autograd bookkeeping; every
FLOP runs in libblindno_b200.so.  There is no CPU implementation: the CPU dispatch key raises.

    differentiable ops (what a model calls)                  reference code it replaces
    ---------------------------------------------------------------------------------------------------
    spectral_conv2d(x, w1, w2, m1, m2, prec)                 SpectralConv2d.forward   2d_FPE/FNOModules.py:156-178
    spectral_conv1d(x, w, m, halve_dc, prec)                 SpectralConv1d.forward   1d_FPE/FNOModules.py:47-59
    fno_lift_pad(x_cl, fc0_w, fc0_b, ndim)                   fc0 + permute + F.pad    FNOModules.py:103-106, :219-224
    fno_layer2d / fno_layer1d(z, w.., conv_w, conv_b, gelu_in, prec)
                                                             layer body               FNOModules.py:226-232, :108-114
    fno_project(z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w) crop + fc1 + GELU + fc2  FNOModules.py:116-121, :234-239
    fno_net(...)                                             FNO1d/FNO2d.forward (+ bag gather / grid concat / bag
                                                             mean + detached lift)    2d_FPE/NIOModules.py:548-575
    bag_pool_lift(s, grid, fc0_w, fc0_b)                     bag mean + detached fc0  2d_FPE/NIOModules.py:564-575
    bag_project_pool_lift(z, n_keep, fc1.., grid, fc0..)     projection of every snapshot, then the above
    deeponet_pool_contract_lift(w, basis, b0, grid, fc0..)   DeepOnetNoBiasOrg.forward + bag mean + lift
                                                             DeepONetModules.py:142-151, 1d_GPE/NIOModules.py:209-219
    bag_attention_mean(x, ln_w, ln_b, eps)                   TemporalSelfAttention.forward + .mean(dim=1)
                                                             2d_FPE/NIOModules.py:1063-1083, :1153-1170
    heads_mse(outs, target)                                  criterion(model(inputs, grid), outputs), MSELoss over the
                                                             concatenated head outputs      2d_FPE/train_fno.py:116,146-147

    *_forward / *_backward ops are the non-differentiable primitives the formulas above are made of.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import FnoParams, FnoShape, LiftInput, PREC_FP32, PREC_TF32, PREC_TF32X3, SpectralShape, check, pad_amount

__all__ = ["FnoSpec", "set_precision", "stage_wfwd", "fno_apply", "spectral_conv", "bag_pool_lift", "adam_step_flat", "heads_mse", "heads_mse_grads", "bag_attention_mean",
           "kernel_launches", "fno_lift_pad", "fno_layer", "fno_project", "OP_NAMES",
           "PREC_FP32", "PREC_TF32", "PREC_TF32X3"]

NS = "blindno_b200"
_LIB = torch.library.Library(NS, "DEF")
_OPS = torch.ops.blindno_b200


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


def _like_param(flat: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """A gradient computed as fp32 (pairs) in the layout of parameter ``ref``."""
    if ref.is_complex():
        return torch.view_as_complex(flat.view(*ref.shape, 2))
    return flat.view(ref.shape)


def _empty(dev):
    return torch.empty(0, dtype=torch.float32, device=dev)


OP_NAMES: List[str] = []


def _define(name: str, schema: str, cuda_impl, fake_impl=None):
    """Schema + CUDA kernel + loud CPU key (+ shape-only fake kernel, also used on the meta device)."""
    _LIB.define(f"{name}{schema}")
    _LIB.impl(name, cuda_impl, "CUDA")

    def _no_cpu(*args, **kwargs):
        raise RuntimeError(f"blindno_b200::{name} runs on CUDA tensors only (there is no CPU fallback by design)")

    _LIB.impl(name, _no_cpu, "CPU")
    if fake_impl is not None:
        torch.library.register_fake(f"{NS}::{name}", fake_impl, lib=_LIB)
    OP_NAMES.append(name)


def _define_composite(name: str, schema: str, fn):
    _LIB.define(f"{name}{schema}")
    _LIB.impl(name, fn, "CompositeImplicitAutograd")
    OP_NAMES.append(name)


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


# ---------------------------------------------------------------------------------------------
# stage_wfwd: one pruned forward DFT (tests, kernel benchmarks)
# ---------------------------------------------------------------------------------------------
def _stage_wfwd_cuda(x, m2, hp=1, m1=0, act=False, prec=0):     # (the dispatcher drops trailing default arguments)
    _need_cuda(x)
    xc = _f32c(x)
    rows, wp = xc.shape
    out = torch.empty(rows, m2, 2, dtype=torch.float32, device=xc.device)
    with torch.cuda.device(xc.device):
        check(_lib.lib().bdn_stage_wfwd(hp, wp, m1, m2, rows, _ptr(xc), _ptr(out), int(act), prec, _stream()),
              "bdn_stage_wfwd")
    return torch.view_as_complex(out)


_define("stage_wfwd", "(Tensor x, int m2, int hp=1, int m1=0, bool act=False, int prec=0) -> Tensor", _stage_wfwd_cuda,
        lambda x, m2, hp=1, m1=0, act=False, prec=0: x.new_empty(x.shape[0], m2, dtype=torch.complex64))


def stage_wfwd(x: torch.Tensor, m2: int, *, hp: int = 1, m1: int = 0, act: bool = False, prec: int = PREC_FP32):
    """One stage on its own (tests, kernel benchmarks): pruned forward DFT along the last axis of
    x [rows, wp] -> complex64 [rows, m2].  ``prec=PREC_TF32`` runs the tcgen05 tensor-core kernel."""
    return _OPS.stage_wfwd(x, m2, hp, m1, act, prec)


# ---------------------------------------------------------------------------------------------
# one spectral convolution
# ---------------------------------------------------------------------------------------------
def _spectral_shape(x_shape, w1_shape, ndim: int, prec: int) -> SpectralShape:
    s = SpectralShape()
    s.ndim, s.prec = ndim, prec
    if ndim == 2:
        s.images, s.c_in, s.hp, s.wp = x_shape
        ci, s.c_out, s.m1, s.m2 = w1_shape[:4]
    else:
        (s.images, s.c_in, s.wp), s.hp, s.m1 = x_shape, 1, 0
        ci, s.c_out, s.m2 = w1_shape[:3]
    if ci != s.c_in:
        raise RuntimeError(f"weights expect {ci} input channels, x has {s.c_in}")
    return s


def _spectral_forward_cuda(x, w1, w2, prec, save):
    L = _lib.lib()
    _need_cuda(x, w1, w2)
    ndim = 2 if w2 is not None else 1
    xc, w1c = _f32c(x), _f32c(w1)
    w2c = _f32c(w2) if w2 is not None else None
    if xc.dim() != ndim + 2:
        raise RuntimeError(f"spectral conv {ndim}-D expects x with {ndim + 2} dims, got {tuple(xc.shape)}")
    s = _spectral_shape(tuple(xc.shape), tuple(w1c.shape), ndim, prec)
    dev = xc.device
    with torch.cuda.device(dev):
        ws_bytes = L.bdn_spectral_workspace_bytes(C.byref(s))
        if ws_bytes == 0:
            check(-1, "bdn_spectral_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        K = 2 * s.m1 if ndim == 2 else 1
        xs = torch.empty(s.images * s.c_in * K * s.m2 * 2, dtype=torch.float32, device=dev) if save else None
        y = torch.empty((s.images, s.c_out) + tuple(xc.shape[2:]), dtype=torch.float32, device=dev)
        check(L.bdn_spectral_forward(C.byref(s), _ptr(xc), _ptr(w1c), _ptr(w2c), _ptr(y), _ptr(xs), _ptr(ws),
                                     ws_bytes, _stream()), "bdn_spectral_forward")
    return y, (xs if xs is not None else _empty(dev))


def _spectral_forward_fake(x, w1, w2, prec, save):
    ndim = 2 if w2 is not None else 1
    c_in, c_out = w1.shape[0], w1.shape[1]
    K = 2 * w1.shape[2] if ndim == 2 else 1
    m2 = w1.shape[3] if ndim == 2 else w1.shape[2]
    y = x.new_empty((x.shape[0], c_out) + tuple(x.shape[2:]), dtype=torch.float32)
    return y, x.new_empty(x.shape[0] * c_in * K * m2 * 2 if save else 0, dtype=torch.float32)


def _spectral_backward_cuda(gy, xs, w1, w2, need_gx, prec):
    L = _lib.lib()
    _need_cuda(gy, xs, w1, w2)
    ndim = 2 if w2 is not None else 1
    gyc, w1c = _f32c(gy), _f32c(w1)
    w2c = _f32c(w2) if w2 is not None else None
    c_in = w1c.shape[0]
    s = _spectral_shape((gyc.shape[0], c_in) + tuple(gyc.shape[2:]), tuple(w1c.shape), ndim, prec)
    dev = gyc.device
    with torch.cuda.device(dev):
        ws_bytes = L.bdn_spectral_workspace_bytes(C.byref(s))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        gx = torch.empty((s.images, c_in) + tuple(gyc.shape[2:]), dtype=torch.float32, device=dev) if need_gx else None
        gw1 = torch.empty_like(w1c)
        gw2 = torch.empty_like(w2c) if w2c is not None else None
        check(L.bdn_spectral_backward(C.byref(s), _ptr(gyc), _ptr(xs), _ptr(w1c), _ptr(w2c), _ptr(gx), _ptr(gw1),
                                      _ptr(gw2), _ptr(ws), ws_bytes, _stream()), "bdn_spectral_backward")
    return (gx if gx is not None else _empty(dev), _like_param(gw1, w1),
            _like_param(gw2, w2) if w2 is not None else _empty(dev))


def _spectral_backward_fake(gy, xs, w1, w2, need_gx, prec):
    gx = gy.new_empty((gy.shape[0], w1.shape[0]) + tuple(gy.shape[2:]) if need_gx else (0,), dtype=torch.float32)
    return gx, torch.empty_like(w1), (torch.empty_like(w2) if w2 is not None else gy.new_empty(0))


_define("spectral_conv_forward", "(Tensor x, Tensor w1, Tensor? w2, int prec, bool save) -> (Tensor, Tensor)",
        _spectral_forward_cuda, _spectral_forward_fake)
_define("spectral_conv_backward",
        "(Tensor gy, Tensor xs, Tensor w1, Tensor? w2, bool need_gx, int prec) -> (Tensor, Tensor, Tensor)",
        _spectral_backward_cuda, _spectral_backward_fake)


def _spectral_setup(ctx, inputs, output):
    x, w1, w2, prec, save = inputs
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(output[1])
    ctx.save_for_backward(output[1], w1, w2)
    ctx.prec, ctx.saved_spectrum = prec, save


def _spectral_bwd(ctx, gy, _gxs):
    xs, w1, w2 = ctx.saved_tensors
    if gy is None:
        return None, None, None, None, None
    if not ctx.saved_spectrum:
        raise RuntimeError("spectral_conv_forward was called with save=False; it cannot be differentiated")
    gx, gw1, gw2 = _OPS.spectral_conv_backward(gy, xs, w1, w2, ctx.needs_input_grad[0], ctx.prec)
    return (gx if ctx.needs_input_grad[0] else None, gw1 if ctx.needs_input_grad[1] else None,
            gw2 if (w2 is not None and ctx.needs_input_grad[2]) else None, None, None)


torch.library.register_autograd(f"{NS}::spectral_conv_forward", _spectral_bwd, setup_context=_spectral_setup, lib=_LIB)


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _spectral_conv2d(x, w1, w2, m1, m2, prec=0):
    if tuple(w1.shape[2:4]) != (m1, m2) or tuple(w2.shape[2:4]) != (m1, m2):
        raise RuntimeError(f"weights keep {tuple(w1.shape[2:4])} modes, the call says ({m1}, {m2})")
    return _OPS.spectral_conv_forward(x, w1, w2, prec, _wants_grad(x, w1, w2))[0]


def _spectral_conv1d(x, w, m, halve_dc=True, prec=0):
    if w.shape[2] != m:
        raise RuntimeError(f"weights keep {w.shape[2]} modes, the call says {m}")
    if not halve_dc:
        raise RuntimeError("the reference's SpectralConv1d halves the DC bin (FNOModules.py:51-52); halve_dc=False is not built")
    return _OPS.spectral_conv_forward(x, w, None, prec, _wants_grad(x, w))[0]


_define_composite("spectral_conv2d", "(Tensor x, Tensor w1, Tensor w2, int m1, int m2, int prec=0) -> Tensor", _spectral_conv2d)
_define_composite("spectral_conv1d", "(Tensor x, Tensor w, int m, bool halve_dc=True, int prec=0) -> Tensor", _spectral_conv1d)


def spectral_conv(x: torch.Tensor, w1: torch.Tensor, w2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SpectralConv2d.forward (x [B,C,H,W], w1/w2 real pairs [...,2] or complex64) or, with ``w2=None``,
    SpectralConv1d.forward (x [B,C,N], w1 complex64 [Ci,Co,m]; DC bin halved as in the reference)."""
    if w2 is not None:
        return _OPS.spectral_conv2d(x, w1, w2, w1.shape[2], w1.shape[3])
    return _OPS.spectral_conv1d(x, w1, w1.shape[2])


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

    def as_ints(self) -> List[int]:
        """The ``int[] spec`` argument of the fno_net ops."""
        return [self.ndim, self.c_in, self.width, self.c_out, self.n_layers, self.modes1, self.modes2, self.hidden, self.prec]

    @staticmethod
    def from_ints(v: Sequence[int]) -> "FnoSpec":
        if len(v) != 9:
            raise RuntimeError("spec must be [ndim, c_in, width, c_out, n_layers, modes1, modes2, hidden, prec]")
        return FnoSpec(*[int(k) for k in v])


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


def _lift_input(spec: FnoSpec, x_cl, bags, idx, grid):
    """-> (LiftInput, keep-alive tensors, images, h, w, n_bags, n_keep, grid_dim, device)"""
    lift = LiftInput()
    if x_cl is not None:
        x_cl = _f32c(x_cl)
        if x_cl.dim() != spec.ndim + 2:
            raise RuntimeError(f"FNO{spec.ndim}d expects a channels-last input with {spec.ndim + 2} dims, got {tuple(x_cl.shape)}")
        if spec.ndim == 2:
            images, h, w, cin = x_cl.shape
        else:
            (images, w, cin), h = x_cl.shape, 1
        if cin != spec.c_in:
            raise RuntimeError(f"input has {cin} features, fc0 expects {spec.c_in}")
        lift.x_cl = _ptr(x_cl)
        return lift, (x_cl,), images, h, w, 0, 0, 0, x_cl.device
    if bags is None or grid is None:
        raise RuntimeError("an FNO net needs x_cl, or bags + grid")
    bags, grid = _f32c(bags), _f32c(grid)
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
    lift.bags, lift.idx, lift.grid = _ptr(bags), _ptr(idx), _ptr(grid)
    lift.n_bags, lift.bag_len, lift.n_keep, lift.grid_dim = n_bags, bag_len, n_keep, gd
    return lift, (bags, idx, grid), n_bags * n_keep, h, w, n_bags, n_keep, gd, bags.device


def _fno_forward_cuda(x_cl, bags, idx, grid, pool_w0, pool_b0, params, spec, save, grad_sink):
    L = _lib.lib()
    spec = FnoSpec.from_ints(spec)
    _need_cuda(x_cl, bags, grid, pool_w0, pool_b0, *params)
    if len(params) != spec.n_params:
        raise RuntimeError(f"expected {spec.n_params} parameter tensors, got {len(params)}")
    flat = [_f32c(p) for p in params]
    lift, keep, images, h, w, n_bags, n_keep, gd, dev = _lift_input(spec, x_cl, bags, idx, grid)
    shape = _make_shape(spec, images, h, w)
    pooled = pool_w0 is not None
    if pooled:
        if x_cl is not None or spec.c_out != 1:
            raise RuntimeError("bag pooling needs the bag input form and a scalar FNO output")
        pool_w0, pool_b0 = _f32c(pool_w0), _f32c(pool_b0)
    with torch.cuda.device(dev):
        ws_bytes = L.bdn_fno_workspace_bytes(C.byref(shape))
        if ws_bytes == 0 and images > 0:
            check(-1, "bdn_fno_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        z_saved = xs_saved = None
        if save:
            z_saved = torch.empty(L.bdn_fno_act_floats(C.byref(shape)), dtype=torch.float32, device=dev)
            xs_saved = torch.empty(L.bdn_fno_spec_floats(C.byref(shape)), dtype=torch.float32, device=dev)
        out_shape = (images, shape.out_h, shape.out_w, spec.c_out) if spec.ndim == 2 else (images, shape.out_w, spec.c_out)
        out = torch.empty(out_shape, dtype=torch.float32, device=dev)
        cparams = _fill_params(spec, flat)
        check(L.bdn_fno_forward(C.byref(shape), C.byref(cparams), C.byref(lift), _ptr(out), _ptr(z_saved),
                                _ptr(xs_saved), _ptr(ws), ws_bytes, _stream()), "bdn_fno_forward")
        result = out
        if pooled:
            npix = shape.out_h * shape.out_w
            if npix != h * w:
                raise RuntimeError("bag pooling needs the FNO output on the input grid (square 2-D grids)")
            width0 = pool_w0.shape[0]
            lifted_shape = (n_bags, h, w, width0) if spec.ndim == 2 else (n_bags, w, width0)
            result = torch.empty(lifted_shape, dtype=torch.float32, device=dev)
            check(L.bdn_bag_pool_lift_forward(_ptr(out), _ptr(keep[2]), _ptr(pool_w0), _ptr(pool_b0), _ptr(result),
                                              n_bags, n_keep, npix, gd, width0, _stream()),
                  "bdn_bag_pool_lift_forward")
    return result, (z_saved if save else _empty(dev)), (xs_saved if save else _empty(dev))


def _fno_dims(x_cl, bags, idx, spec: FnoSpec):
    if x_cl is not None:
        images = x_cl.shape[0]
        h, w = (x_cl.shape[1], x_cl.shape[2]) if spec.ndim == 2 else (1, x_cl.shape[1])
        return images, h, w, 0
    n_keep = idx.numel() if idx is not None else bags.shape[1]
    h, w = (bags.shape[2], bags.shape[3]) if spec.ndim == 2 else (1, bags.shape[2])
    return bags.shape[0] * n_keep, h, w, bags.shape[0]


def _fno_forward_fake(x_cl, bags, idx, grid, pool_w0, pool_b0, params, spec, save, grad_sink):
    spec = FnoSpec.from_ints(spec)
    images, h, w, n_bags = _fno_dims(x_cl, bags, idx, spec)
    s = _make_shape(spec, images, h, w)
    ref = x_cl if x_cl is not None else bags
    if pool_w0 is not None:
        shp = (n_bags, h, w, pool_w0.shape[0]) if spec.ndim == 2 else (n_bags, w, pool_w0.shape[0])
    else:
        shp = (images, s.out_h, s.out_w, spec.c_out) if spec.ndim == 2 else (images, s.out_w, spec.c_out)
    L = _lib.lib()          # the size queries are host-only
    nz = L.bdn_fno_act_floats(C.byref(s)) if save else 0
    nx = L.bdn_fno_spec_floats(C.byref(s)) if save else 0
    return (ref.new_empty(shp, dtype=torch.float32), ref.new_empty(nz, dtype=torch.float32),
            ref.new_empty(nx, dtype=torch.float32))


def _fno_backward_cuda(g, x_cl, bags, idx, grid, pool_w0, params, z_saved, xs_saved, spec, need_gx, grad_sink):
    L = _lib.lib()
    spec = FnoSpec.from_ints(spec)
    _need_cuda(g, z_saved, xs_saved, *params)
    flat = [_f32c(p) for p in params]
    lift, keep, images, h, w, n_bags, n_keep, gd, dev = _lift_input(spec, x_cl, bags, idx, grid)
    shape = _make_shape(spec, images, h, w)
    pooled = pool_w0 is not None
    g = _f32c(g)
    with torch.cuda.device(dev):
        if pooled:
            pool_w0 = _f32c(pool_w0)
            npix = shape.out_h * shape.out_w
            gpool = torch.empty(n_bags * npix, dtype=torch.float32, device=dev)
            check(L.bdn_bag_pool_lift_backward(_ptr(g), _ptr(pool_w0), _ptr(gpool), n_bags, npix, gd,
                                               pool_w0.shape[0], _stream()), "bdn_bag_pool_lift_backward")
            g = gpool
        sizes = [t.numel() for t in flat]
        offs, total = slot_layout(sizes)
        if grad_sink is not None:
            # the trainer's flat gradient buffer: kernels accumulate into it, autograd sees no grads
            if grad_sink.numel() != total or grad_sink.dtype != torch.float32 or not grad_sink.is_contiguous():
                raise RuntimeError("gradient sink does not match this net's slot layout")
            gflat = grad_sink
        else:
            gflat = torch.zeros(total, dtype=torch.float32, device=dev)
        gviews = [gflat[o:o + n] for o, n in zip(offs, sizes)]
        cgrads = _fill_params(spec, gviews)
        gx = torch.empty_like(keep[0]) if (x_cl is not None and need_gx) else None
        ws_bytes = L.bdn_fno_workspace_bytes(C.byref(shape))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        cparams = _fill_params(spec, flat)
        check(L.bdn_fno_backward(C.byref(shape), C.byref(cparams), C.byref(lift), _ptr(g), int(pooled),
                                 max(n_keep, 1), _ptr(z_saved), _ptr(xs_saved), C.byref(cgrads), _ptr(gx),
                                 _ptr(ws), ws_bytes, _stream()), "bdn_fno_backward")
    return (gx if gx is not None else _empty(dev)), (gflat if grad_sink is None else _empty(dev))


def _fno_backward_fake(g, x_cl, bags, idx, grid, pool_w0, params, z_saved, xs_saved, spec, need_gx, grad_sink):
    sizes = [p.numel() * (2 if p.is_complex() else 1) for p in params]
    _, total = slot_layout(sizes)
    gx = torch.empty_like(x_cl, dtype=torch.float32) if (x_cl is not None and need_gx) else g.new_empty(0)
    return gx, g.new_empty(total if grad_sink is None else 0, dtype=torch.float32)


_define("fno_net_forward",
        "(Tensor? x_cl, Tensor? bags, Tensor? idx, Tensor? grid, Tensor? pool_w0, Tensor? pool_b0, Tensor[] params, "
        "int[] spec, bool save, Tensor? grad_sink) -> (Tensor, Tensor, Tensor)", _fno_forward_cuda, _fno_forward_fake)
_define("fno_net_backward",
        "(Tensor g, Tensor? x_cl, Tensor? bags, Tensor? idx, Tensor? grid, Tensor? pool_w0, Tensor[] params, "
        "Tensor z_saved, Tensor xs_saved, int[] spec, bool need_gx, Tensor(a!)? grad_sink) -> (Tensor, Tensor)",
        _fno_backward_cuda, _fno_backward_fake)


def _fno_setup(ctx, inputs, output):
    x_cl, bags, idx, grid, pool_w0, pool_b0, params, spec, save, grad_sink = inputs
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(output[1], output[2])
    ctx.n_params = len(params)
    ctx.save_for_backward(output[1], output[2], x_cl, bags, idx, grid, pool_w0, *params)
    ctx.spec, ctx.saved_acts = list(spec), save
    ctx.sink = grad_sink      # mutated by the backward kernels: deliberately not a saved (version-checked) tensor


def _fno_bwd(ctx, g, _gz, _gxs):
    none = (None, None, None, None, None, None, [None] * ctx.n_params, None, None, None)
    if g is None:
        return none
    if not ctx.saved_acts:
        raise RuntimeError("fno_net_forward was called with save=False; it cannot be differentiated")
    z_saved, xs_saved, x_cl, bags, idx, grid, pool_w0, *params = ctx.saved_tensors
    need = ctx.needs_input_grad
    need_gx = bool(x_cl is not None and need[0])
    gx, gflat = _OPS.fno_net_backward(g, x_cl, bags, idx, grid, pool_w0, params, z_saved, xs_saved, ctx.spec, need_gx, ctx.sink)
    grads: List[Optional[torch.Tensor]] = [None] * ctx.n_params
    if ctx.sink is None:
        sizes = [p.numel() * (2 if p.is_complex() else 1) for p in params]
        offs, _ = slot_layout(sizes)
        for k, (p, o, n) in enumerate(zip(params, offs, sizes)):
            if need[6][k]:
                grads[k] = _like_param(gflat[o:o + n], p)
    return (gx if need_gx else None, None, None, None, None, None, grads, None, None, None)


torch.library.register_autograd(f"{NS}::fno_net_forward", _fno_bwd, setup_context=_fno_setup, lib=_LIB)


def _fno_net(x_cl, bags, idx, grid, pool_w0, pool_b0, params, spec, grad_sink=None):
    save = _wants_grad(x_cl, *params)
    return _OPS.fno_net_forward(x_cl, bags, idx, grid, pool_w0, pool_b0, params, spec, save, grad_sink)[0]


_define_composite("fno_net",
                  "(Tensor? x_cl, Tensor? bags, Tensor? idx, Tensor? grid, Tensor? pool_w0, Tensor? pool_b0, "
                  "Tensor[] params, int[] spec, Tensor? grad_sink=None) -> Tensor", _fno_net)


def fno_apply(spec: FnoSpec, params: Sequence[torch.Tensor], *, x_cl=None, bags=None, idx=None, grid=None,
              pool=None, grad_sink=None) -> torch.Tensor:
    """Run one FNO net (``torch.ops.blindno_b200.fno_net``).

    ``x_cl``: channels-last input [images, (h,) w, c_in]; or ``bags`` [B, L0, (h,) w] + ``grid`` [(h,) w, d]
    (+ optional int ``idx`` of kept snapshots) for the per-snapshot NIO-FNO encoder, whose input
    concat(snapshot, grid) is never materialised.  ``pool=(fc0.weight, fc0.bias)`` additionally applies the
    bag mean and the detached lift, returning [B, (h,) w, width] instead of the per-snapshot outputs.
    ``grad_sink``: a flat fp32 buffer laid out by ``slot_layout`` over ``params``; when given, the backward
    kernels accumulate the parameter gradients straight into it (the data-parallel trainer all-reduces that
    buffer in one call) and autograd receives no parameter gradients.
    """
    pw, pb = pool if pool is not None else (None, None)
    return _OPS.fno_net(x_cl, bags, idx, grid, pw, pb, list(params), spec.as_ints(), grad_sink)


# ---------------------------------------------------------------------------------------------
# single stages: lift, layer body, projection
# ---------------------------------------------------------------------------------------------
def _stage_shape(ndim, images, *, c_in=1, width=1, c_out=1, hidden=1, h=1, w=1, hp=1, wp=1, out_h=1, out_w=1, m1=None,
                 m2=1, prec=0) -> FnoShape:
    s = FnoShape()
    s.ndim, s.images, s.c_in, s.width, s.c_out, s.hidden, s.n_layers = ndim, images, c_in, width, c_out, hidden, 1
    s.h, s.w, s.hp, s.wp, s.out_h, s.out_w = h, w, hp, wp, out_h, out_w
    s.m1 = (1 if ndim == 2 else 0) if m1 is None else m1
    s.m2, s.prec = m2, prec
    return s


def _lift_shape(x_cl, fc0_w, ndim):
    if ndim not in (1, 2) or x_cl.dim() != ndim + 2:
        raise RuntimeError(f"fno_lift_pad: ndim={ndim} needs a channels-last input with {ndim + 2} dims, got {tuple(x_cl.shape)}")
    if ndim == 2:
        images, h, w, cin = x_cl.shape
        hp = h + pad_amount(h)
    else:
        (images, w, cin), h, hp = x_cl.shape, 1, 1
    width = fc0_w.shape[0]
    if fc0_w.shape[1] != cin:
        raise RuntimeError(f"input has {cin} features, fc0 expects {fc0_w.shape[1]}")
    wp = w + pad_amount(w)
    return images, h, w, hp, wp, cin, width


def _lift_cuda(x_cl, fc0_w, fc0_b, ndim):
    _need_cuda(x_cl, fc0_w, fc0_b)
    x, w0, b0 = _f32c(x_cl), _f32c(fc0_w), _f32c(fc0_b)
    images, h, w, hp, wp, cin, width = _lift_shape(x, w0, ndim)
    s = _stage_shape(ndim, images, c_in=cin, width=width, h=h, w=w, hp=hp, wp=wp, out_h=hp, out_w=wp)
    lift = LiftInput()
    lift.x_cl = _ptr(x)
    z0 = torch.empty((images, width, hp, wp) if ndim == 2 else (images, width, wp), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().bdn_stage_lift_forward(C.byref(s), _ptr(w0), _ptr(b0), C.byref(lift), _ptr(z0), _stream()),
              "bdn_stage_lift_forward")
    return z0


def _lift_fake(x_cl, fc0_w, fc0_b, ndim):
    images, h, w, hp, wp, cin, width = _lift_shape(x_cl, fc0_w, ndim)
    return x_cl.new_empty((images, width, hp, wp) if ndim == 2 else (images, width, wp), dtype=torch.float32)


def _lift_backward_cuda(gz0, x_cl, fc0_w, fc0_b, ndim, need_gx):
    _need_cuda(gz0, x_cl, fc0_w, fc0_b)
    g, x, w0, b0 = _f32c(gz0), _f32c(x_cl), _f32c(fc0_w), _f32c(fc0_b)
    images, h, w, hp, wp, cin, width = _lift_shape(x, w0, ndim)
    s = _stage_shape(ndim, images, c_in=cin, width=width, h=h, w=w, hp=hp, wp=wp, out_h=hp, out_w=wp)
    lift = LiftInput()
    lift.x_cl = _ptr(x)
    gw, gb = torch.zeros_like(w0), torch.zeros_like(b0)
    gx = torch.empty_like(x) if need_gx else None
    with torch.cuda.device(x.device):
        check(_lib.lib().bdn_stage_lift_backward(C.byref(s), _ptr(w0), _ptr(b0), C.byref(lift), _ptr(g), _ptr(gw), _ptr(gb),
                                                 _ptr(gx), _stream()), "bdn_stage_lift_backward")
    return (gx if gx is not None else _empty(x.device)), gw, gb


_define("fno_lift_pad", "(Tensor x_cl, Tensor fc0_w, Tensor fc0_b, int ndim) -> Tensor", _lift_cuda, _lift_fake)
_define("fno_lift_pad_backward",
        "(Tensor gz0, Tensor x_cl, Tensor fc0_w, Tensor fc0_b, int ndim, bool need_gx) -> (Tensor, Tensor, Tensor)",
        _lift_backward_cuda,
        lambda gz0, x_cl, fc0_w, fc0_b, ndim, need_gx: (torch.empty_like(x_cl) if need_gx else gz0.new_empty(0),
                                                       torch.empty_like(fc0_w), torch.empty_like(fc0_b)))


def _lift_setup(ctx, inputs, output):
    x_cl, fc0_w, fc0_b, ndim = inputs
    ctx.save_for_backward(x_cl, fc0_w, fc0_b)
    ctx.ndim = ndim


def _lift_bwd(ctx, gz0):
    x_cl, fc0_w, fc0_b = ctx.saved_tensors
    need = ctx.needs_input_grad
    gx, gw, gb = _OPS.fno_lift_pad_backward(gz0, x_cl, fc0_w, fc0_b, ctx.ndim, need[0])
    return gx if need[0] else None, gw if need[1] else None, gb if need[2] else None, None


torch.library.register_autograd(f"{NS}::fno_lift_pad", _lift_bwd, setup_context=_lift_setup, lib=_LIB)


def fno_lift_pad(x_cl, fc0_w, fc0_b, ndim: int):
    """fc0 + channels-first + zero pad round(n/4): x_cl [images,(h,)w,c_in] -> z0 [images,width,(hp,)wp]."""
    return _OPS.fno_lift_pad(x_cl, fc0_w, fc0_b, ndim)


def _layer_shape(z, w1, w2, prec):
    ndim = 2 if w2 is not None else 1
    if z.dim() != ndim + 2:
        raise RuntimeError(f"fno_layer{ndim}d expects z with {ndim + 2} dims, got {tuple(z.shape)}")
    images, width = z.shape[0], z.shape[1]
    hp, wp = (z.shape[2], z.shape[3]) if ndim == 2 else (1, z.shape[2])
    if w1.shape[0] != width or w1.shape[1] != width:
        raise RuntimeError(f"spectral weights are {tuple(w1.shape[:2])}, z has {width} channels")
    m1, m2 = (w1.shape[2], w1.shape[3]) if ndim == 2 else (0, w1.shape[2])
    s = _stage_shape(ndim, images, c_in=width, width=width, h=hp, w=wp, hp=hp, wp=wp, out_h=hp, out_w=wp, m1=m1, m2=m2, prec=prec)
    K = 2 * m1 if ndim == 2 else 1
    return s, images * width * K * m2 * 2


def _layer_forward_cuda(z, w1, w2, conv_w, conv_b, gelu_in, prec, save):
    L = _lib.lib()
    _need_cuda(z, w1, w2, conv_w, conv_b)
    zc, w1c, cw, cb = _f32c(z), _f32c(w1), _f32c(conv_w), _f32c(conv_b)
    w2c = _f32c(w2) if w2 is not None else None
    s, nxs = _layer_shape(zc, w1c, w2c, prec)
    dev = zc.device
    with torch.cuda.device(dev):
        ws_bytes = L.bdn_stage_layer_workspace_bytes(C.byref(s))
        if ws_bytes == 0:
            check(-1, "bdn_stage_layer_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        xs = torch.empty(nxs, dtype=torch.float32, device=dev) if save else None
        out = torch.empty_like(zc)
        check(L.bdn_stage_layer_forward(C.byref(s), _ptr(zc), int(gelu_in), _ptr(w1c), _ptr(w2c), _ptr(cw), _ptr(cb),
                                        _ptr(out), _ptr(xs), _ptr(ws), ws_bytes, _stream()), "bdn_stage_layer_forward")
    return out, (xs if save else _empty(dev))


def _layer_forward_fake(z, w1, w2, conv_w, conv_b, gelu_in, prec, save):
    _, nxs = _layer_shape(z, torch.view_as_real(w1) if w1.is_complex() else w1, w2, prec)
    return torch.empty_like(z), z.new_empty(nxs if save else 0)


def _layer_backward_cuda(gz_out, z, xs, w1, w2, conv_w, gelu_in, prec):
    L = _lib.lib()
    _need_cuda(gz_out, z, xs, w1, w2, conv_w)
    g, zc, w1c, cw = _f32c(gz_out), _f32c(z), _f32c(w1), _f32c(conv_w)
    w2c = _f32c(w2) if w2 is not None else None
    s, _ = _layer_shape(zc, w1c, w2c, prec)
    dev = zc.device
    with torch.cuda.device(dev):
        ws_bytes = L.bdn_stage_layer_workspace_bytes(C.byref(s))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        gz_in = torch.empty_like(zc)
        gw1 = torch.zeros_like(w1c)
        gw2 = torch.zeros_like(w2c) if w2c is not None else None
        gcw = torch.zeros_like(cw)
        gcb = torch.zeros(cw.shape[0], dtype=torch.float32, device=dev)
        check(L.bdn_stage_layer_backward(C.byref(s), _ptr(g), _ptr(zc), int(gelu_in), _ptr(xs), _ptr(w1c), _ptr(w2c),
                                         _ptr(cw), _ptr(gz_in), _ptr(gw1), _ptr(gw2), _ptr(gcw), _ptr(gcb), _ptr(ws),
                                         ws_bytes, _stream()), "bdn_stage_layer_backward")
    return (gz_in, _like_param(gw1, w1), _like_param(gw2, w2) if w2 is not None else _empty(dev),
            gcw.view(conv_w.shape), gcb)


_define("fno_layer_forward",
        "(Tensor z, Tensor w1, Tensor? w2, Tensor conv_w, Tensor conv_b, bool gelu_in, int prec, bool save) -> (Tensor, Tensor)",
        _layer_forward_cuda, _layer_forward_fake)
_define("fno_layer_backward",
        "(Tensor gz_out, Tensor z, Tensor xs, Tensor w1, Tensor? w2, Tensor conv_w, bool gelu_in, int prec) -> "
        "(Tensor, Tensor, Tensor, Tensor, Tensor)", _layer_backward_cuda,
        lambda gz_out, z, xs, w1, w2, conv_w, gelu_in, prec: (
            torch.empty_like(z), torch.empty_like(w1), torch.empty_like(w2) if w2 is not None else z.new_empty(0),
            torch.empty_like(conv_w), z.new_empty(conv_w.shape[0])))


def _layer_setup(ctx, inputs, output):
    z, w1, w2, conv_w, conv_b, gelu_in, prec, save = inputs
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(output[1])
    ctx.save_for_backward(z, output[1], w1, w2, conv_w)
    ctx.gelu_in, ctx.prec, ctx.saved_spectrum = gelu_in, prec, save


def _layer_bwd(ctx, gz_out, _gxs):
    if gz_out is None:
        return (None,) * 8
    if not ctx.saved_spectrum:
        raise RuntimeError("fno_layer_forward was called with save=False; it cannot be differentiated")
    z, xs, w1, w2, conv_w = ctx.saved_tensors
    gz, gw1, gw2, gcw, gcb = _OPS.fno_layer_backward(gz_out, z, xs, w1, w2, conv_w, ctx.gelu_in, ctx.prec)
    need = ctx.needs_input_grad
    return (gz if need[0] else None, gw1 if need[1] else None, gw2 if (w2 is not None and need[2]) else None,
            gcw if need[3] else None, gcb if need[4] else None, None, None, None)


torch.library.register_autograd(f"{NS}::fno_layer_forward", _layer_bwd, setup_context=_layer_setup, lib=_LIB)


def _fno_layer2d(z, w1, w2, conv_w, conv_b, gelu_in, prec=0):
    return _OPS.fno_layer_forward(z, w1, w2, conv_w, conv_b, gelu_in, prec, _wants_grad(z, w1, w2, conv_w, conv_b))[0]


def _fno_layer1d(z, w, conv_w, conv_b, gelu_in, prec=0):
    return _OPS.fno_layer_forward(z, w, None, conv_w, conv_b, gelu_in, prec, _wants_grad(z, w, conv_w, conv_b))[0]


_define_composite("fno_layer2d", "(Tensor z, Tensor w1, Tensor w2, Tensor conv_w, Tensor conv_b, bool gelu_in, int prec=0) -> Tensor",
                  _fno_layer2d)
_define_composite("fno_layer1d", "(Tensor z, Tensor w, Tensor conv_w, Tensor conv_b, bool gelu_in, int prec=0) -> Tensor",
                  _fno_layer1d)


def fno_layer(z, w1, w2, conv_w, conv_b, gelu_in: bool, prec: int = PREC_FP32):
    """One layer body on pre-activation tensors: ``spectral(a) + conv1x1(a) + b`` with ``a = gelu(z)`` if
    ``gelu_in`` else ``z`` (the reference's ``x = gelu(spectral(x) + conv(x))`` with the GELU moved to the
    consumer; no GELU follows the last layer)."""
    if w2 is not None:
        return _OPS.fno_layer2d(z, w1, w2, conv_w, conv_b, gelu_in, prec)
    return _OPS.fno_layer1d(z, w1, conv_w, conv_b, gelu_in, prec)


def _project_shape(z, fc1_w, fc2_w, out_h, out_w):
    ndim = z.dim() - 2
    if ndim not in (1, 2):
        raise RuntimeError(f"fno_project expects z [images, width, (hp,) wp], got {tuple(z.shape)}")
    images, width = z.shape[0], z.shape[1]
    hp, wp = (z.shape[2], z.shape[3]) if ndim == 2 else (1, z.shape[2])
    if ndim == 1 and out_h != 1:
        raise RuntimeError("1-D projection needs out_h = 1")
    if fc1_w.shape[1] != width or fc2_w.shape[1] != fc1_w.shape[0]:
        raise RuntimeError("fc1 / fc2 shapes do not match z")
    s = _stage_shape(ndim, images, c_in=width, width=width, c_out=fc2_w.shape[0], hidden=fc1_w.shape[0], h=hp, w=wp, hp=hp,
                     wp=wp, out_h=out_h, out_w=out_w)
    out_shape = (images, out_h, out_w, fc2_w.shape[0]) if ndim == 2 else (images, out_w, fc2_w.shape[0])
    return s, out_shape


def _project_cuda(z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w):
    _need_cuda(z, fc1_w, fc1_b, fc2_w, fc2_b)
    zc, w1, b1, w2, b2 = (_f32c(t) for t in (z, fc1_w, fc1_b, fc2_w, fc2_b))
    s, out_shape = _project_shape(zc, w1, w2, out_h, out_w)
    out = torch.empty(out_shape, dtype=torch.float32, device=zc.device)
    with torch.cuda.device(zc.device):
        check(_lib.lib().bdn_stage_project_forward(C.byref(s), _ptr(zc), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(out),
                                                   _stream()), "bdn_stage_project_forward")
    return out


def _project_backward_cuda(g, z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w, n_keep):
    _need_cuda(g, z, fc1_w, fc1_b, fc2_w, fc2_b)
    gc, zc, w1, b1, w2, b2 = (_f32c(t) for t in (g, z, fc1_w, fc1_b, fc2_w, fc2_b))
    s, out_shape = _project_shape(zc, w1, w2, out_h, out_w)
    pooled = n_keep > 0          # g is the gradient of the bag mean: [images / n_keep, ...]
    gz = torch.empty_like(zc)
    gw1, gb1, gw2, gb2 = (torch.zeros_like(t) for t in (w1, b1, w2, b2))
    with torch.cuda.device(zc.device):
        check(_lib.lib().bdn_stage_project_backward(C.byref(s), _ptr(zc), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(gc),
                                                    int(pooled), max(n_keep, 1), _ptr(gz), _ptr(gw1), _ptr(gb1), _ptr(gw2),
                                                    _ptr(gb2), _stream()), "bdn_stage_project_backward")
    return gz, gw1, gb1, gw2, gb2


_define("fno_project", "(Tensor z, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, int out_h, int out_w) -> Tensor",
        _project_cuda,
        lambda z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w: z.new_empty(_project_shape(z, fc1_w, fc2_w, out_h, out_w)[1]))
_define("fno_project_backward",
        "(Tensor g, Tensor z, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, int out_h, int out_w, int n_keep) -> "
        "(Tensor, Tensor, Tensor, Tensor, Tensor)", _project_backward_cuda,
        lambda g, z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w, n_keep: tuple(
            torch.empty_like(t) for t in (z, fc1_w, fc1_b, fc2_w, fc2_b)))


def _project_setup(ctx, inputs, output):
    z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w = inputs
    ctx.save_for_backward(z, fc1_w, fc1_b, fc2_w, fc2_b)
    ctx.crop = (out_h, out_w)


def _project_bwd(ctx, g):
    z, fc1_w, fc1_b, fc2_w, fc2_b = ctx.saved_tensors
    grads = _OPS.fno_project_backward(g, z, fc1_w, fc1_b, fc2_w, fc2_b, ctx.crop[0], ctx.crop[1], 0)
    return (*[gr if need else None for gr, need in zip(grads, ctx.needs_input_grad[:5])], None, None)


torch.library.register_autograd(f"{NS}::fno_project", _project_bwd, setup_context=_project_setup, lib=_LIB)


def fno_project(z, fc1_w, fc1_b, fc2_w, fc2_b, out_h: int, out_w: int):
    """crop + fc1 + exact GELU + fc2: z [images,width,(hp,)wp] -> [images,(out_h,)out_w,c_out]."""
    return _OPS.fno_project(z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w)


# ---------------------------------------------------------------------------------------------
# bag mean + detached lift (the NIO models feed it DeepONet outputs), and the two pooled tails
# ---------------------------------------------------------------------------------------------
def _pool_dims(s, grid, w0):
    n_bags, n_keep = s.shape[:2]
    gshape = tuple(s.shape[2:])
    npix = 1
    for d in gshape:
        npix *= d
    return n_bags, n_keep, gshape, npix, grid.shape[-1], w0.shape[0]


def _pool_lift_cuda(s, grid, w0, b0):
    _need_cuda(s, grid, w0, b0)
    sc, gc, w0c, b0c = _f32c(s), _f32c(grid), _f32c(w0), _f32c(b0)
    n_bags, n_keep, gshape, npix, gd, width = _pool_dims(sc, gc, w0c)
    out = torch.empty((n_bags,) + gshape + (width,), dtype=torch.float32, device=sc.device)
    with torch.cuda.device(sc.device):
        check(_lib.lib().bdn_bag_pool_lift_forward(_ptr(sc), _ptr(gc), _ptr(w0c), _ptr(b0c), _ptr(out), n_bags, n_keep,
                                                   npix, gd, width, _stream()), "bdn_bag_pool_lift_forward")
    return out


def _pool_lift_backward_cuda(g, w0, grid_dim):
    _need_cuda(g, w0)
    gc, w0c = _f32c(g), _f32c(w0)
    n_bags, width = gc.shape[0], w0c.shape[0]
    npix = gc.numel() // max(n_bags * width, 1)
    gpool = torch.empty(n_bags, npix, dtype=torch.float32, device=gc.device)
    with torch.cuda.device(gc.device):
        check(_lib.lib().bdn_bag_pool_lift_backward(_ptr(gc), _ptr(w0c), _ptr(gpool), n_bags, npix, grid_dim, width,
                                                    _stream()), "bdn_bag_pool_lift_backward")
    return gpool


_define("bag_pool_lift", "(Tensor s, Tensor grid, Tensor fc0_w, Tensor fc0_b) -> Tensor", _pool_lift_cuda,
        lambda s, grid, w0, b0: s.new_empty((s.shape[0],) + tuple(s.shape[2:]) + (w0.shape[0],)))
_define("bag_pool_lift_backward", "(Tensor g, Tensor fc0_w, int grid_dim) -> Tensor", _pool_lift_backward_cuda,
        lambda g, w0, grid_dim: g.new_empty(g.shape[0], g.numel() // max(g.shape[0] * w0.shape[0], 1)))


def _pool_setup(ctx, inputs, output):
    s, grid, w0, b0 = inputs
    ctx.save_for_backward(w0)
    ctx.s_shape, ctx.grid_dim = tuple(s.shape), grid.shape[-1]


def _pool_bwd(ctx, g):
    (w0,) = ctx.saved_tensors
    if not ctx.needs_input_grad[0]:
        return None, None, None, None
    n_bags, n_keep = ctx.s_shape[:2]
    gpool = _OPS.bag_pool_lift_backward(g, w0, ctx.grid_dim)
    gs = (gpool.view(n_bags, 1, -1) / n_keep).expand(n_bags, n_keep, gpool.shape[1]).reshape(ctx.s_shape)
    return gs, None, None, None       # fc0 is detached in the reference (.data): no gradient, by construction


torch.library.register_autograd(f"{NS}::bag_pool_lift", _pool_bwd, setup_context=_pool_setup, lib=_LIB)


def bag_pool_lift(s, grid, w0, b0):
    """fc0([grid, mean_l s_l]) with fc0 detached: s [B,L,*g], grid [*g,d] -> [B,*g,width]."""
    return _OPS.bag_pool_lift(s, grid, w0, b0)


def _bag_project_pool_lift(z, n_keep, fc1_w, fc1_b, fc2_w, fc2_b, grid, fc0_w, fc0_b):
    """Projection of every snapshot's last activation, bag mean, detached lift (NIO-FNO tail)."""
    gshape = tuple(grid.shape[:-1])
    out_h, out_w = (gshape if len(gshape) == 2 else (1, gshape[0]))
    s = _OPS.fno_project(z, fc1_w, fc1_b, fc2_w, fc2_b, out_h, out_w)          # [B*L, *g, 1]
    return _OPS.bag_pool_lift(s.reshape(z.shape[0] // n_keep, n_keep, *gshape), grid, fc0_w.detach(), fc0_b.detach())


def _nio_tail_dims(w, basis, grid, fc0_w):
    if w.dim() != 3 or basis.dim() != 2 or basis.shape[1] != w.shape[2]:
        raise RuntimeError(f"nio_tail expects w [bags, keep, p] and basis [points, p], got {tuple(w.shape)}, {tuple(basis.shape)}")
    gshape = tuple(grid.shape[:-1])
    npix = 1
    for d in gshape:
        npix *= d
    if npix != basis.shape[0] or fc0_w.shape[1] != grid.shape[-1] + 1:
        raise RuntimeError("nio_tail: grid / basis / fc0 shapes do not match")
    return w.shape[0], w.shape[1], w.shape[2], gshape, npix, grid.shape[-1], fc0_w.shape[0]


def _nio_tail_forward_cuda(w, basis, b0, grid, fc0_w, fc0_b):
    _need_cuda(w, basis, b0, grid, fc0_w, fc0_b)
    wc, bc, b0c, gc, w0c, fbc = (_f32c(t) for t in (w, basis, b0, grid, fc0_w, fc0_b))
    n_bags, n_keep, p, gshape, npix, gd, width = _nio_tail_dims(wc, bc, gc, w0c)
    out = torch.empty((n_bags,) + gshape + (width,), dtype=torch.float32, device=wc.device)
    wbar = torch.empty(n_bags, p, dtype=torch.float32, device=wc.device)
    with torch.cuda.device(wc.device):
        check(_lib.lib().bdn_nio_tail_forward(_ptr(wc), _ptr(bc), _ptr(b0c), _ptr(gc), _ptr(w0c), _ptr(fbc), _ptr(out),
                                              _ptr(wbar), n_bags, n_keep, p, npix, gd, width, _stream()), "bdn_nio_tail_forward")
    return out, wbar


def _nio_tail_backward_cuda(g, basis, wbar, fc0_w, n_keep, grid_dim):
    _need_cuda(g, basis, wbar, fc0_w)
    gc, bc, wbc, w0c = (_f32c(t) for t in (g, basis, wbar, fc0_w))
    n_bags, p = wbc.shape
    npix, width = bc.shape[0], w0c.shape[0]
    g_w = torch.empty(n_bags, n_keep, p, dtype=torch.float32, device=gc.device)
    g_basis = torch.empty_like(bc)
    g_b0 = torch.empty(1, dtype=torch.float32, device=gc.device)
    scratch = torch.empty(n_bags, p, dtype=torch.float32, device=gc.device)
    with torch.cuda.device(gc.device):
        check(_lib.lib().bdn_nio_tail_backward(_ptr(gc), _ptr(bc), _ptr(wbc), _ptr(w0c), _ptr(g_w), _ptr(g_basis), _ptr(g_b0),
                                               _ptr(scratch), n_bags, n_keep, p, npix, grid_dim, width, _stream()),
              "bdn_nio_tail_backward")
    return g_w, g_basis, g_b0


def _nio_tail_fake(w, basis, b0, grid, fc0_w, fc0_b):
    n_bags, _, p, gshape, _, _, width = _nio_tail_dims(w, basis, grid, fc0_w)
    return w.new_empty((n_bags,) + gshape + (width,)), w.new_empty(n_bags, p)


_define("nio_tail_forward", "(Tensor w, Tensor basis, Tensor b0, Tensor grid, Tensor fc0_w, Tensor fc0_b) -> (Tensor, Tensor)",
        _nio_tail_forward_cuda, _nio_tail_fake)
_define("nio_tail_backward", "(Tensor g, Tensor basis, Tensor wbar, Tensor fc0_w, int n_keep, int grid_dim) -> "
        "(Tensor, Tensor, Tensor)", _nio_tail_backward_cuda,
        lambda g, basis, wbar, fc0_w, n_keep, grid_dim: (g.new_empty(wbar.shape[0], n_keep, wbar.shape[1]),
                                                         torch.empty_like(basis), g.new_empty(1)))


def _nio_tail_setup(ctx, inputs, output):
    w, basis, b0, grid, fc0_w, fc0_b = inputs
    ctx.save_for_backward(basis, output[1], fc0_w)
    ctx.n_keep, ctx.grid_dim, ctx.b0_shape = w.shape[1], grid.shape[-1], tuple(b0.shape)
    ctx.mark_non_differentiable(output[1])


def _nio_tail_bwd(ctx, g, _g_wbar):
    basis, wbar, fc0_w = ctx.saved_tensors
    need = ctx.needs_input_grad
    if not any(need[:3]):
        return (None,) * 6
    g_w, g_basis, g_b0 = _OPS.nio_tail_backward(g, basis, wbar, fc0_w, ctx.n_keep, ctx.grid_dim)
    # fc0 is detached in the reference (.data): no gradient, by construction
    return (g_w if need[0] else None, g_basis if need[1] else None, g_b0.reshape(ctx.b0_shape) if need[2] else None,
            None, None, None)


torch.library.register_autograd(f"{NS}::nio_tail_forward", _nio_tail_bwd, setup_context=_nio_tail_setup, lib=_LIB)


def _deeponet_pool_contract_lift(w, basis, b0, grid, fc0_w, fc0_b):
    """((mean_l w_l) @ basis^T + b0) / sqrt(p), then the detached lift, in one kernel: by linearity the bag mean is taken
    on the [B, L, p] branch coefficients, so the [B, L, n_points] DeepONet output is never materialised (NIO tail, K6)."""
    return _OPS.nio_tail_forward(w, basis, b0, grid, fc0_w.detach(), fc0_b.detach())[0]


_define_composite("bag_project_pool_lift",
                  "(Tensor z, int n_keep, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, Tensor grid, Tensor fc0_w, "
                  "Tensor fc0_b) -> Tensor", _bag_project_pool_lift)
_define_composite("deeponet_pool_contract_lift",
                  "(Tensor w, Tensor basis, Tensor b0, Tensor grid, Tensor fc0_w, Tensor fc0_b) -> Tensor",
                  _deeponet_pool_contract_lift)


# ---------------------------------------------------------------------------------------------
# bag attention + bag mean of the BlinDNO models (no [L, D] intermediates)
# ---------------------------------------------------------------------------------------------
BAG_ATTENTION_MAX_KEEP = 128


def _bag_attention_dims(x, ln_w, ln_b):
    if x.dim() != 3:
        raise RuntimeError(f"bag_attention_mean expects tokens [bags, keep, dim], got {tuple(x.shape)}")
    n_bags, n_keep, dim = x.shape
    if ln_w.shape != (dim,) or ln_b.shape != (dim,):
        raise RuntimeError("bag_attention_mean: LayerNorm weight / bias must have one entry per token feature")
    if n_keep > BAG_ATTENTION_MAX_KEEP:
        raise RuntimeError(f"bag_attention_mean: bags of more than {BAG_ATTENTION_MAX_KEEP} snapshots are not built")
    return n_bags, n_keep, dim


def _bag_attention_forward_cuda(x, ln_w, ln_b, eps):
    _need_cuda(x, ln_w, ln_b)
    xc, wc, bc = _f32c(x), _f32c(ln_w), _f32c(ln_b)
    n_bags, n_keep, dim = _bag_attention_dims(xc, wc, bc)
    L = _lib.lib()
    out = torch.empty(n_bags, dim, dtype=torch.float32, device=xc.device)
    saved = torch.empty(L.bdn_bag_attention_saved_floats(max(n_bags, 1), n_keep), dtype=torch.float32, device=xc.device)
    with torch.cuda.device(xc.device):
        check(L.bdn_bag_attention_mean_forward(_ptr(xc), _ptr(wc), _ptr(bc), _ptr(out), _ptr(saved), n_bags, n_keep, dim,
                                               float(eps), _stream()), "bdn_bag_attention_mean_forward")
    return out, saved


def _bag_attention_backward_cuda(g, x, ln_w, saved):
    _need_cuda(g, x, ln_w, saved)
    gc, xc, wc = _f32c(g), _f32c(x), _f32c(ln_w)
    n_bags, n_keep, dim = xc.shape
    L = _lib.lib()
    gx = torch.empty_like(xc)
    gw_bag = torch.empty(n_bags, dim, dtype=torch.float32, device=xc.device)
    ws = torch.empty(L.bdn_bag_attention_workspace_floats(max(n_bags, 1), n_keep), dtype=torch.float32, device=xc.device)
    with torch.cuda.device(xc.device):
        check(L.bdn_bag_attention_mean_backward(_ptr(xc), _ptr(gc), _ptr(wc), _ptr(saved), _ptr(gx), _ptr(gw_bag), _ptr(ws),
                                                n_bags, n_keep, dim, _stream()), "bdn_bag_attention_mean_backward")
    return gx, gw_bag


def _bag_attention_fake(x, ln_w, ln_b, eps):
    n_bags, n_keep, dim = _bag_attention_dims(x, ln_w, ln_b)
    return x.new_empty(n_bags, dim), x.new_empty(n_bags * (4 * n_keep + 4 + 2 * n_keep * n_keep))


_define("bag_attention_mean_forward", "(Tensor x, Tensor ln_w, Tensor ln_b, float eps) -> (Tensor, Tensor)",
        _bag_attention_forward_cuda, _bag_attention_fake)
_define("bag_attention_mean_backward", "(Tensor g, Tensor x, Tensor ln_w, Tensor saved) -> (Tensor, Tensor)",
        _bag_attention_backward_cuda, lambda g, x, ln_w, saved: (torch.empty_like(x), x.new_empty(x.shape[0], x.shape[2])))


def _bag_attention_setup(ctx, inputs, output):
    x, ln_w, ln_b, eps = inputs
    ctx.save_for_backward(x, ln_w, output[1])
    ctx.mark_non_differentiable(output[1])


def _bag_attention_bwd(ctx, g, _g_saved):
    x, ln_w, saved = ctx.saved_tensors
    gx, gw_bag = _OPS.bag_attention_mean_backward(g, x, ln_w, saved)
    need = ctx.needs_input_grad
    return (gx if need[0] else None, gw_bag.sum(dim=0) if need[1] else None, g.sum(dim=0) if need[2] else None, None)


torch.library.register_autograd(f"{NS}::bag_attention_mean_forward", _bag_attention_bwd, setup_context=_bag_attention_setup,
                                lib=_LIB)


def bag_attention_mean(x, ln_w, ln_b, eps: float = 1e-5):
    """``LayerNorm(softmax(x x^T / sqrt(D)) x + x).mean(dim=1)`` for tokens x [bags, keep, D] (TemporalSelfAttention followed
    by the bag mean, 2d_FPE/NIOModules.py:1063-1083, :1153-1170) -> [bags, D]; everything between the two passes over x is
    keep x keep algebra (csrc/bagattn.cu)."""
    return _OPS.bag_attention_mean_forward(x, ln_w, ln_b, eps)[0]


# ---------------------------------------------------------------------------------------------
# training loss: MSE over the (never formed) concatenation of the head outputs
# ---------------------------------------------------------------------------------------------
_MSE_SCRATCH = {}


def _mse_scratch(dev):
    """65 floats per (device, stream): the forward kernel's block counter (left at zero by every launch) and partials."""
    key = (dev.index, _stream())
    buf = _MSE_SCRATCH.get(key)
    if buf is None:
        buf = _MSE_SCRATCH[key] = torch.zeros(65, dtype=torch.float32, device=dev)
    return buf


def _mse_dims(outs, target):
    n = len(outs)
    if not 1 <= n <= 4:
        raise RuntimeError(f"heads_mse takes 1 to 4 head outputs, got {n}")
    c = outs[0].shape[-1]
    npix = outs[0].numel() // max(c, 1)
    for o in outs:
        if o.shape != outs[0].shape:
            raise RuntimeError("heads_mse: head outputs must have one shape")
    if tuple(target.shape) != tuple(outs[0].shape[:-1]) + (n * c,):
        raise RuntimeError(f"heads_mse: target {tuple(target.shape)} does not match {n} heads of {tuple(outs[0].shape)}")
    return n, c, npix


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _heads_mse_launch(outs, target, with_grads):
    _need_cuda(*outs, target)
    oc, tc = [_f32c(o) for o in outs], _f32c(target)
    n, c, npix = _mse_dims(oc, tc)
    loss = torch.empty((), dtype=torch.float32, device=tc.device)
    gs = [torch.empty_like(o) for o in oc] if with_grads else []
    with torch.cuda.device(tc.device):
        check(_lib.lib().bdn_mse_heads_forward(_ptr_array(oc), n, c, npix, _ptr(tc), _ptr(loss),
                                               _ptr_array(gs) if with_grads else None, _ptr(_mse_scratch(tc.device)),
                                               _stream()), "bdn_mse_heads_forward")
    return loss, gs


def _heads_mse_forward_cuda(outs, target):
    return _heads_mse_launch(outs, target, False)[0]


def _heads_mse_grads_cuda(outs, target):
    return _heads_mse_launch(outs, target, True)


def _heads_mse_backward_cuda(outs, target, grad_loss):
    _need_cuda(*outs, target, grad_loss)
    oc, tc, gl = [_f32c(o) for o in outs], _f32c(target), _f32c(grad_loss)
    n, c, npix = _mse_dims(oc, tc)
    gs = [torch.empty_like(o) for o in oc]
    with torch.cuda.device(tc.device):
        check(_lib.lib().bdn_mse_heads_backward(_ptr_array(oc), n, c, npix, _ptr(tc), _ptr(gl), _ptr_array(gs), _stream()),
              "bdn_mse_heads_backward")
    return gs


_define("heads_mse", "(Tensor[] outs, Tensor target) -> Tensor", _heads_mse_forward_cuda,
        lambda outs, target: (_mse_dims(outs, target), target.new_empty(()))[1])
_define("heads_mse_grads", "(Tensor[] outs, Tensor target) -> (Tensor, Tensor[])", _heads_mse_grads_cuda,
        lambda outs, target: ((_mse_dims(outs, target), target.new_empty(()))[1], [torch.empty_like(o) for o in outs]))
_define("heads_mse_backward", "(Tensor[] outs, Tensor target, Tensor grad_loss) -> Tensor[]", _heads_mse_backward_cuda,
        lambda outs, target, grad_loss: [torch.empty_like(o) for o in outs])


def _mse_setup(ctx, inputs, output):
    outs, target = inputs
    ctx.save_for_backward(target, *outs)


def _mse_bwd(ctx, g):
    target, *outs = ctx.saved_tensors
    return _OPS.heads_mse_backward(list(outs), target, g), None      # the target gets no gradient


torch.library.register_autograd(f"{NS}::heads_mse", _mse_bwd, setup_context=_mse_setup, lib=_LIB)


def heads_mse_grads(outs, target):
    """(loss, [d loss / d out_k]) of ``F.mse_loss(torch.cat(outs, -1), target)`` in ONE kernel: what a train step needs
    (its backward starts from grad_loss = 1); not differentiable -- feed the gradients to ``torch.autograd.backward``."""
    loss, gs = _OPS.heads_mse_grads([o.detach() for o in outs], target)
    return loss, list(gs)


def heads_mse(outs, target):
    """``F.mse_loss(torch.cat(outs, -1), target)`` in one kernel (and one for its backward); the sum is deterministic."""
    return _OPS.heads_mse(list(outs), target)


# ---------------------------------------------------------------------------------------------
# fused Adam over a flat buffer
# ---------------------------------------------------------------------------------------------
def _adam_cuda(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale):
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    with torch.cuda.device(param.device):
        check(_lib.lib().bdn_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(),
                                       lr, beta1, beta2, eps, step, grad_scale, _stream()), "bdn_adam_step")


_define("adam_step_flat_",
        "(Tensor(a!) param, Tensor grad, Tensor(b!) exp_avg, Tensor(c!) exp_avg_sq, float lr, float beta1, float beta2, "
        "float eps, int step, float grad_scale) -> ()", _adam_cuda,
        lambda param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale: None)


def adam_step_flat(param, grad, exp_avg, exp_avg_sq, *, lr, betas=(0.9, 0.999), eps=1e-8, step, grad_scale=1.0):
    """torch.optim.Adam's update over one flat fp32 buffer, in place, one kernel."""
    _OPS.adam_step_flat_(param, grad, exp_avg, exp_avg_sq, lr, betas[0], betas[1], eps, step, grad_scale)
