"""GPU (-m gpu): the torch.library custom ops, one fused stage at a time, against the CPU oracle.

The stage ops (fno_lift_pad -> fno_layer{1,2}d x n -> fno_project) chained by hand must reproduce both
the oracle's FNO forward/backward and the whole-net op (which runs the same kernels back to back).
Tolerances as in tests/test_gpu_parity.py: 1e-5 relative on outputs; gradients within
max(1e-5 * scale, 3 x the reference's own fp32 error) of the fp64 oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from blindno_b200 import ops
from blindno_b200.surface import fno
from oracle import blindno_oracle as O
from tests.helpers import rel_err
from tests.test_gpu_parity import _gmax, _grad_check, _leaf, _to64

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"
NSO = torch.ops.blindno_b200


def _stage_chain(net, x, ndim):
    z = ops.fno_lift_pad(x, net.fc0.weight, net.fc0.bias, ndim)
    for k in range(net.n_layers):
        s = net.spectral_list[k]
        z = ops.fno_layer(z, s.weights1, getattr(s, "weights2", None), net.conv_list[k].weight, net.conv_list[k].bias, k > 0)
    if ndim == 2:
        h, w = x.shape[1], x.shape[2]
        out_h, out_w = z.shape[2] - O.pad_amount(w), z.shape[3] - O.pad_amount(h)
    else:
        out_h, out_w = 1, x.shape[1]
    return ops.fno_project(z, net.fc1.weight, net.fc1.bias, net.fc2.weight, net.fc2.bias, out_h, out_w)


@pytest.mark.parametrize("ndim,shape,ctor", [
    (2, (3, 21, 17, 3), dict(modes=5, width=6, n_layers=3, input_dim=3, output_dim=1)),
    (2, (2, 61, 61, 12), dict(modes=32, width=12, n_layers=2, input_dim=12, output_dim=1)),
    (1, (5, 80, 2), dict(modes=15, width=30, n_layers=3, input_dim=2, output_dim=2)),
    (1, (4, 33, 3), dict(modes=7, width=5, n_layers=2, input_dim=3, output_dim=1)),
    (1, (3, 128, 20), dict(modes=40, width=20, n_layers=3, input_dim=20, output_dim=1)),     # the 1D-GPE head (fused 1-D layer)
    (1, (2, 800, 2), dict(modes=64, width=6, n_layers=2, input_dim=2, output_dim=1)),        # too wide for the fused 1-D layer
    (1, (3, 40, 32), dict(modes=20, width=32, n_layers=8, input_dim=32, output_dim=4)),      # the C ABI's limits: width, c_in, layers, c_out
    (2, (2, 20, 24, 32), dict(modes=6, width=32, n_layers=2, input_dim=32, output_dim=1)),
])
def test_stage_ops_chain_vs_oracle_and_whole_net(ndim, shape, ctor):
    torch.manual_seed(3)
    net = (fno.FNO2d if ndim == 2 else fno.FNO1d)(**ctor)
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(DEV)
    x = torch.randn(*shape)
    oracle = O.fno2d_forward if ndim == 2 else O.fno1d_forward

    xg = x.to(DEV).requires_grad_(True)
    got = _stage_chain(net, xg, ndim)
    gy = torch.randn(got.shape, generator=torch.Generator().manual_seed(5))
    got.backward(gy.to(DEV))
    stage_grads = {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}
    stage_gx = xg.grad.detach().cpu().clone()

    refs = []
    for cast in (lambda d: d, _to64):
        p = _leaf(cast(dict(params)))
        xx = (x.double() if cast is _to64 else x).clone().requires_grad_(True)
        y = oracle(p, xx)
        y.backward(gy.double() if cast is _to64 else gy)
        refs.append((y.detach(), {k: v.grad for k, v in p.items()}, xx.grad))
    (y32, g32, gx32), (y64, g64, gx64) = refs
    assert rel_err(got, y64) < TOL
    floor = 1.2e-7 * max(_gmax(g64), gx64.abs().max().item())
    for k in stage_grads:
        _grad_check(k, stage_grads[k], g32[k], g64[k], floor=floor)
    _grad_check("x", stage_gx, gx32, gx64, floor=floor)

    # the whole-net op runs the same kernels: same forward up to summation order, gradients up to atomics order
    net.zero_grad()
    xw = x.to(DEV).requires_grad_(True)
    whole = net(xw)
    assert rel_err(whole, got) < 5e-6      # (few-image nets read a mode-major weight copy: other summation order)
    whole.backward(gy.to(DEV))
    for k, p in net.named_parameters():
        _grad_check("whole/" + k, p.grad, g32[k], g64[k], floor=floor)


def test_layer_op_alone_vs_oracle_including_gelu_on_load():
    torch.manual_seed(11)
    C, hp, wp, m = 6, 20, 24, 4
    z = torch.randn(3, C, hp, wp)
    w1, w2 = torch.rand(C, C, m, m, 2) / C, torch.rand(C, C, m, m, 2) / C
    cw, cb = torch.randn(C, C, 1, 1) / C, torch.randn(C)
    for gelu_in in (False, True):
        args = [t.to(DEV).requires_grad_(True) for t in (z, w1, w2, cw, cb)]
        out = NSO.fno_layer2d(*args, gelu_in)
        gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
        out.backward(gy.to(DEV))
        ref = [t.double().requires_grad_(True) for t in (z, w1, w2, cw, cb)]
        a = F.gelu(ref[0]) if gelu_in else ref[0]
        want = O.spectral_conv2d(a, ref[1], ref[2]) + F.conv2d(a, ref[3], ref[4])
        want.backward(gy.double())
        assert rel_err(out, want) < TOL
        for got_t, ref_t, name in zip(args, ref, ("z", "w1", "w2", "conv_w", "conv_b")):
            assert rel_err(got_t.grad, ref_t.grad) < 2e-5, (name, gelu_in)


@pytest.mark.parametrize("hidden,c_out,width,shape", [(128, 1, 12, (4, 12, 20, 24)), (256, 2, 30, (6, 30, 50)), (512, 3, 30, (6, 30, 50)),
                                                    (512, 1, 4, (300, 4, 40, 40))])
def test_project_op_other_hidden_sizes(hidden, c_out, width, shape):
    """fno_project on its own with fc1 widths other than the reference's 128 (the C ABI serves up to 512): few-pixel
    and many-pixel regimes, including the size at which the backward's warp-private accumulators no longer fit
    shared memory and the shuffle + shared-atomic reduction is used instead."""
    torch.manual_seed(hidden + c_out)
    nd = len(shape) - 2
    z = torch.randn(*shape)
    fc1_w, fc1_b = torch.randn(hidden, width) / width ** 0.5, torch.randn(hidden)
    fc2_w, fc2_b = torch.randn(c_out, hidden) / hidden ** 0.5, torch.randn(c_out)
    out_h, out_w = (shape[2] - 3, shape[3] - 4) if nd == 2 else (1, shape[2] - 7)
    args = [t.to(DEV).requires_grad_(True) for t in (z, fc1_w, fc1_b, fc2_w, fc2_b)]
    out = NSO.fno_project(*args, out_h, out_w)
    gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    out.backward(gy.to(DEV))
    ref = [t.double().requires_grad_(True) for t in (z, fc1_w, fc1_b, fc2_w, fc2_b)]
    crop = ref[0][..., :out_h, :out_w].permute(0, 2, 3, 1) if nd == 2 else ref[0][..., :out_w].transpose(1, 2)
    want = F.linear(F.gelu(F.linear(crop, ref[1], ref[2])), ref[3], ref[4])
    want.backward(gy.double())
    assert rel_err(out, want) < TOL
    for a, r, name in zip(args, ref, ("z", "fc1_w", "fc1_b", "fc2_w", "fc2_b")):
        assert rel_err(a.grad, r.grad) < 2e-5, name


def test_pooled_tails_vs_oracle():
    torch.manual_seed(7)
    B, L, n, C = 2, 5, 12, 4
    pad = O.pad_amount(n)
    ax = torch.linspace(-1, 1, n)
    grid = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), dim=-1)
    z = torch.randn(B * L, C, n + pad, n + pad)
    fc1_w, fc1_b, fc2_w, fc2_b = torch.randn(128, C) / 2, torch.randn(128), torch.randn(1, 128) / 11, torch.randn(1)
    fc0_w, fc0_b = torch.randn(7, 3), torch.randn(7)

    dev_args = [t.to(DEV).requires_grad_(True) for t in (z, fc1_w, fc1_b, fc2_w, fc2_b)]
    got = NSO.bag_project_pool_lift(dev_args[0], L, *dev_args[1:], grid.to(DEV), fc0_w.to(DEV), fc0_b.to(DEV))
    gy = torch.randn(got.shape, generator=torch.Generator().manual_seed(1))
    got.backward(gy.to(DEV))
    ref = [t.double().requires_grad_(True) for t in (z, fc1_w, fc1_b, fc2_w, fc2_b)]
    h = ref[0][..., :n, :n].permute(0, 2, 3, 1)
    s = F.linear(F.gelu(F.linear(h, ref[1], ref[2])), ref[3], ref[4]).reshape(B, L, n, n)
    want = O.bag_pool_lift(s, grid.double(), fc0_w.double(), fc0_b.double())
    want.backward(gy.double())
    assert rel_err(got, want) < TOL
    for a, r in zip(dev_args, ref):
        assert rel_err(a.grad, r.grad) < 2e-5

    w = torch.randn(B, L, 25)
    basis = torch.randn(n * n, 25)
    b0 = torch.tensor(0.3)
    dev_args = [t.to(DEV).requires_grad_(True) for t in (w, basis, b0)]
    got = NSO.deeponet_pool_contract_lift(*dev_args, grid.to(DEV), fc0_w.to(DEV), fc0_b.to(DEV))
    got.backward(gy.to(DEV))
    ref = [t.double().requires_grad_(True) for t in (w, basis, b0)]
    per_snapshot = O.deeponet_forward({"deeponet.b0": ref[2]}, ref[0], ref[1]).reshape(B, L, n, n)   # the reference order
    want = O.bag_pool_lift(per_snapshot, grid.double(), fc0_w.double(), fc0_b.double())
    want.backward(gy.double())
    assert rel_err(got, want) < TOL
    for a, r in zip(dev_args, ref):
        assert rel_err(a.grad, r.grad) < 1e-4      # TF32-free cuBLAS GEMM vs fp64; cancelling sums over n*n points


@pytest.mark.parametrize("n_bags,n_keep,p,grid_shape", [(4, 75, 25, (128,)), (2, 9, 130, (11, 7)), (3, 100, 64, (61, 61))])
def test_nio_tail_other_basis_sizes_and_grids(n_bags, n_keep, p, grid_shape):
    """deeponet_pool_contract_lift (the one-kernel NIO tail) in 1-D and 2-D, with more basis functions than threads of a
    block column (p = 130), against DeepOnetNoBiasOrg.forward + bag mean + detached lift in fp64."""
    g = torch.Generator().manual_seed(p)
    npix = int(np.prod(grid_shape))
    gd = len(grid_shape)
    w, basis, b0 = torch.randn(n_bags, n_keep, p, generator=g), torch.randn(npix, p, generator=g), torch.tensor(-0.4)
    grid = torch.rand(*grid_shape, gd, generator=g)
    fc0_w, fc0_b = torch.randn(6, gd + 1, generator=g), torch.randn(6, generator=g)
    gy = torch.randn(n_bags, *grid_shape, 6, generator=g)
    dev = [t.to(DEV).requires_grad_(True) for t in (w, basis, b0)]
    got = NSO.deeponet_pool_contract_lift(*dev, grid.to(DEV), fc0_w.to(DEV), fc0_b.to(DEV))
    got.backward(gy.to(DEV))
    ref = [t.double().requires_grad_(True) for t in (w, basis, b0)]
    s = O.deeponet_forward({"deeponet.b0": ref[2]}, ref[0], ref[1]).reshape(n_bags, n_keep, *grid_shape)
    want = O.bag_pool_lift(s, grid.double(), fc0_w.double(), fc0_b.double())
    want.backward(gy.double())
    assert rel_err(got, want) < TOL
    for a, r in zip(dev, ref):
        assert rel_err(a.grad, r.grad) < 2e-5


@pytest.mark.parametrize("n_heads,c,shape", [(2, 1, (4, 61, 61)), (1, 1, (4, 128)), (2, 2, (3, 7, 5)), (3, 1, (40, 80, 80))])
def test_heads_mse_matches_mse_loss_on_the_concatenation(n_heads, c, shape):
    """criterion(model(inputs, grid), outputs) with MSELoss (2d_FPE/train_fno.py:116,146-147) without the torch.cat:
    loss and the gradient of every head output, against F.mse_loss in fp64; deterministic across repeated launches."""
    g = torch.Generator().manual_seed(n_heads * 10 + c)
    outs = [torch.randn(*shape, c, generator=g) for _ in range(n_heads)]
    target = torch.randn(*shape, n_heads * c, generator=g)
    dev = [o.to(DEV).requires_grad_(True) for o in outs]
    loss = ops.heads_mse(dev, target.to(DEV))
    (3.0 * loss).backward()
    ref = [o.double().requires_grad_(True) for o in outs]
    want = F.mse_loss(torch.cat(ref, dim=-1), target.double())
    (3.0 * want).backward()
    assert abs(loss.item() - want.item()) <= 2e-6 * abs(want.item())
    for a, r in zip(dev, ref):
        assert rel_err(a.grad, r.grad) < TOL
    again = [ops.heads_mse([d.detach() for d in dev], target.to(DEV)).item() for _ in range(3)]
    assert all(v == loss.item() for v in again)
    # the train step's form: loss and d loss / d outs (grad_loss = 1) from one launch
    loss1, grads1 = ops.heads_mse_grads(dev, target.to(DEV))
    assert loss1.item() == loss.item()
    for gk, r in zip(grads1, ref):
        assert rel_err(gk, r.grad / 3.0) < TOL


@pytest.mark.parametrize("n_bags,n_keep,dim,shift", [(4, 75, 3721, 0.0), (2, 100, 900, 0.5), (3, 51, 144, 0.0), (1, 7, 37, 2.0),
                                                     (2, 128, 392, 0.0)])
def test_bag_attention_mean_vs_the_reference_formulation(n_bags, n_keep, dim, shift):
    """TemporalSelfAttention + .mean(dim=1) (2d_FPE/NIOModules.py:1063-1083, :1153-1170) as the reference writes it, in
    fp64, against the fused op: outputs and the gradients of the tokens and of the LayerNorm affine.  ``shift`` moves
    the token means away from zero (the row variances come from the CENTERED Gram matrix)."""
    g = torch.Generator().manual_seed(n_keep * 7 + dim)
    x = torch.randn(n_bags, n_keep, dim, generator=g) * 0.7 + shift
    ln_w = 1.0 + 0.3 * torch.randn(dim, generator=g)
    ln_b = 0.2 * torch.randn(dim, generator=g)
    gy = torch.randn(n_bags, dim, generator=g)
    dev = [t.to(DEV).requires_grad_(True) for t in (x, ln_w, ln_b)]
    got = ops.bag_attention_mean(*dev, 1e-5)
    got.backward(gy.to(DEV))
    ref = [t.double().requires_grad_(True) for t in (x, ln_w, ln_b)]
    xf = ref[0]
    attn = torch.softmax(torch.matmul(xf, xf.transpose(1, 2)) / dim ** 0.5, dim=-1)
    want = F.layer_norm(torch.matmul(attn, xf) + xf, (dim,), ref[1], ref[2], 1e-5).mean(dim=1)
    want.backward(gy.double())
    assert rel_err(got, want) < TOL
    for a, r in zip(dev, ref):
        assert rel_err(a.grad, r.grad) < 2e-5


def test_opcheck_schema_fake_and_autograd_registration():
    x = torch.randn(2, 3, 8, 8, device=DEV, requires_grad=True)
    w = (torch.rand(3, 3, 2, 2, 2, device=DEV) / 9).requires_grad_(True)
    utils = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(NSO.spectral_conv_forward.default, (x, w, w, 0, True), test_utils=utils)
    xin = torch.randn(2, 8, 8, 3, device=DEV, requires_grad=True)
    fc0_w = torch.randn(4, 3, device=DEV, requires_grad=True)
    fc0_b = torch.randn(4, device=DEV, requires_grad=True)
    torch.library.opcheck(NSO.fno_lift_pad.default, (xin, fc0_w, fc0_b, 2), test_utils=utils)
    z = torch.randn(2, 4, 10, 10, device=DEV, requires_grad=True)
    fc1 = torch.nn.Linear(4, 128).to(DEV)
    fc2 = torch.nn.Linear(128, 1).to(DEV)
    torch.library.opcheck(NSO.fno_project.default, (z, fc1.weight, fc1.bias, fc2.weight, fc2.bias, 8, 8), test_utils=utils)
    net = fno.FNO2d(3, 4, 2, 3, 1).to(DEV)
    torch.library.opcheck(NSO.fno_net_forward.default,
                          (xin, None, None, None, None, None, net._params(), net._spec().as_ints(), True, None),
                          test_utils=utils)
    p = torch.randn(64, device=DEV)
    torch.library.opcheck(NSO.adam_step_flat_.default, (p, torch.randn_like(p), torch.zeros_like(p), torch.zeros_like(p),
                                                        1e-3, 0.9, 0.999, 1e-8, 1, 1.0), test_utils=("test_schema",))
