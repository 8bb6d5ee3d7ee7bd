"""GPU (-m gpu): the fused tensor-core layer kernels (csrc/tc_layer.cu: tc_p = W + H forward DFT, tc_q = channel mix +
H + W inverse DFT + layer epilogue; all four GEMMs on tcgen05) against the CPU oracle's layer body
(2d_FPE/FNOModules.py:156-178, 226-232), one layer at a time through the C ABI's stage entry points.

Bounds: 3xTF32 (BDN_PREC_TF32X3) meets the fp32 rule of tests/test_gpu_parity.py -- outputs 1e-5 relative, gradients
within max(1e-5 * scale, 3 x the reference's own fp32 error, fp32 epsilon x the largest gradient); plain TF32 the stated
2e-3 (outputs) / 1e-2 (gradients).  The profile tags prove which kernels ran: a tensor-core mode that cannot be served
says so (bdn_fno_layer_path) instead of silently computing in another arithmetic."""
import ctypes

import pytest
import torch

from blindno_b200 import _lib, ops
from oracle import blindno_oracle as O
from tests.helpers import rel_err
from tests.test_gpu_parity import _grad_check

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5
TF32_TOL, TF32_GRAD_TOL = 2e-3, 1e-2

CASES = [
    # images, width, hp, wp, m1, m2
    (8, 4, 76, 76, 12, 12),        # few images: one channel plane per CTA
    (150, 4, 76, 76, 12, 12),      # the per-snapshot net: whole images per CTA
    (4, 12, 76, 76, 32, 32),       # the output heads: inputs streamed through the staging buffers
    (37, 4, 76, 76, 12, 12),       # a ragged last wave
    (160, 8, 40, 52, 6, 10),       # rectangular, 4 of 8 channels per CTA
    (3, 6, 28, 36, 5, 7),          # odd mode counts: K padding, scalar store path
    (150, 4, 100, 100, 12, 12),    # 2D-NC grid (80 -> 100)
]


def _layer_ref(z, w1, w2, cw, cb, gelu_in):
    x = torch.nn.functional.gelu(z) if gelu_in else z
    y = O.spectral_conv2d(x, w1, w2)
    return y + torch.einsum("oi,bihw->bohw", cw, x) + cb[None, :, None, None]


def _inputs(case, seed=0):
    images, c, hp, wp, m1, m2 = case
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(images, c, hp, wp, generator=g)
    w1 = torch.rand(c, c, m1, m2, 2, generator=g) / (c * c)
    w2 = torch.rand(c, c, m1, m2, 2, generator=g) / (c * c)
    cw = torch.randn(c, c, generator=g) / c ** 0.5
    cb = torch.randn(c, generator=g)
    gy = torch.randn(images, c, hp, wp, generator=g)
    return z, w1, w2, cw, cb, gy


def _run_gpu(case, prec, gelu_in):
    z, w1, w2, cw, cb, gy = _inputs(case)
    c = case[1]
    leaves = [t.to(DEV).requires_grad_(True) for t in (z, w1, w2, cw, cb)]
    ops.profile_begin()
    out = ops.fno_layer(leaves[0], leaves[1], leaves[2], leaves[3].view(c, c, 1, 1), leaves[4], gelu_in, prec)
    out.backward(gy.to(DEV))
    torch.cuda.synchronize()
    tags = sorted(ops.profile_end())
    return [out.detach().cpu()] + [t.grad.detach().cpu() for t in leaves], tags


def _run_ref(case, gelu_in, dtype):
    z, w1, w2, cw, cb, gy = _inputs(case)
    leaves = [t.to(dtype).requires_grad_(True) for t in (z, w1, w2, cw, cb)]
    out = _layer_ref(*leaves, gelu_in)
    out.backward(gy.to(dtype))
    return [out.detach()] + [t.grad for t in leaves]


def _path(case, prec):
    images, c, hp, wp, m1, m2 = case
    s = ops._stage_shape(2, images, c_in=c, width=c, h=hp, w=wp, hp=hp, wp=wp, out_h=hp, out_w=wp, m1=m1, m2=m2, prec=prec)
    return _lib.lib().bdn_fno_layer_path(ctypes.byref(s))


NAMES = ["z_out", "gz_in", "g_weights1", "g_weights2", "g_conv_w", "g_conv_b"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("gelu_in", [False, True])
def test_layer_3xtf32_meets_the_fp32_bound(case, gelu_in):
    assert _path(case, ops.PREC_FP32) == 0
    path = _path(case, ops.PREC_TF32X3)
    got, tags = _run_gpu(case, ops.PREC_TF32X3, gelu_in)
    if path == 1:
        assert {"tc_p", "tc_q_fwd", "tc_q_bwd"} <= {t.split("/")[0].replace("_gelu", "") for t in tags}, tags
        assert not any(t.startswith(("wfwd", "core2d", "winv")) for t in tags), tags
    else:                       # the shape does not fit: FFMA layer kernels, and the path query said so beforehand
        assert not any(t.startswith("tc_") for t in tags), tags
    ref32, ref64 = _run_ref(case, gelu_in, torch.float32), _run_ref(case, gelu_in, torch.float64)
    assert rel_err(got[0], ref64[0]) < TOL
    gmax = max(float(t.abs().max()) for t in ref64[1:])
    for name, g, r32, r64 in zip(NAMES[1:], got[1:], ref32[1:], ref64[1:]):
        _grad_check(name, g.reshape(r64.shape), r32, r64, floor=1.2e-7 * gmax)


@pytest.mark.parametrize("case", [CASES[1], CASES[2], CASES[5]])
def test_layer_tf32_within_the_stated_bound_and_really_tf32(case):
    assert _path(case, ops.PREC_TF32) == 1
    got, tags = _run_gpu(case, ops.PREC_TF32, True)
    assert any(t.startswith("tc_q_fwd") for t in tags), tags
    ref64 = _run_ref(case, True, torch.float64)
    assert rel_err(got[0], ref64[0]) < TF32_TOL
    errs = [rel_err(g.reshape(r.shape), r) for g, r in zip(got[1:], ref64[1:])]
    assert max(errs) < TF32_GRAD_TOL, dict(zip(NAMES[1:], errs))
    assert max(errs) > 2e-5, "one MMA per K step cannot be this exact: the 3xTF32 path must have run instead"


def test_tensor_core_layer_agrees_with_the_ffma_layer_bitwise_independent_paths():
    """Same inputs through the two independent implementations (FFMA kernels of spectral.cu, tcgen05 kernels of
    tc_layer.cu): they agree to fp32 rounding, far inside either one's distance to the oracle's bound."""
    case = CASES[1]
    a, _ = _run_gpu(case, ops.PREC_FP32, True)
    b, _ = _run_gpu(case, ops.PREC_TF32X3, True)
    for name, x, y in zip(NAMES, a, b):
        assert rel_err(y, x) < 3e-6, name


def test_unsupported_shape_is_reported_not_hidden():
    """A grid beyond the tensor-core kernels' shared-memory plan (padded 160 x 160): the path query answers 0, the
    layer still computes (FFMA kernels, the W-forward stage alone on tcgen05) and meets the bound."""
    case = (3, 4, 160, 160, 12, 12)
    assert _path(case, ops.PREC_TF32X3) == 0
    got, tags = _run_gpu(case, ops.PREC_TF32X3, True)
    assert not any(t.startswith(("tc_p", "tc_q")) for t in tags), tags
    assert any(t.startswith("core2d") for t in tags), tags
    ref64 = _run_ref(case, True, torch.float64)
    assert rel_err(got[0], ref64[0]) < TOL


def test_stage_wfwd_tensor_core_request_runs_on_tensor_cores_or_fails():
    """bdn_stage_wfwd in a tensor-core mode either launches the tcgen05 kernel (profile tag wfwd_tc*) or returns
    BDN_ERR_UNSUPPORTED -- it never computes the stage in another arithmetic behind the caller's back.  Shapes of the
    configs[4] sweep: padded widths 160 and 320, 32 and 64 modes."""
    g = torch.Generator().manual_seed(0)
    ran, refused = [], []
    for wp, m2 in [(76, 12), (160, 32), (160, 64), (320, 32), (320, 64)]:
        x = torch.randn(4096, wp, generator=g).to(DEV)
        for prec in (ops.PREC_TF32, ops.PREC_TF32X3):
            ops.profile_begin()
            try:
                got = ops.stage_wfwd(x, m2, hp=1, m1=0, prec=prec)
                torch.cuda.synchronize()
            except _lib.BlindnoError as e:
                ops.profile_end()
                assert "does not fit the tcgen05 kernel" in str(e)
                refused.append((wp, m2, prec))
                continue
            tags = sorted(ops.profile_end())
            assert tags and all(t.startswith("wfwd_tc") for t in tags), (wp, m2, prec, tags)
            want = torch.fft.rfft(x.cpu().double(), dim=1)[:, :m2]
            want[:, 0] *= 1.0          # (the DC halving of the 1-D layer is applied by the mix stage, not here)
            assert rel_err(torch.view_as_real(got.cpu()), torch.view_as_real(want)) < (TOL if prec == ops.PREC_TF32X3 else TF32_TOL)
            ran.append((wp, m2, prec))
    assert (76, 12, ops.PREC_TF32X3) in ran and (160, 32, ops.PREC_TF32) in ran
    assert len(ran) + len(refused) == 10
