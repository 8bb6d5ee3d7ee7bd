"""CPU: the oracle restatement against golden vectors produced by the reference itself."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import Fixture, assert_grads_close, rel_err
from oracle import blindno_oracle as O
from oracle import dft64

TOL = 2e-6  # fp32 CPU vs fp32 CPU, same library: only summation-order noise


def _leaf(params):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() or v.is_complex() else v)
            for k, v in params.items()}


def _check(fx, fn, *extra, **kw):
    p = _leaf(fx.params)
    x = fx.t("x").clone().requires_grad_(True)
    y = fn(p, x, *extra, **kw)
    assert rel_err(y, fx.t("y")) < TOL
    y.backward(fx.t("gy"))
    assert rel_err(x.grad, fx.t("gx")) < 5 * TOL
    assert_grads_close({k: v.grad for k, v in p.items()}, fx.grads, 5e-5)  # fp32 summation-order noise on cancelling sums reaches 3e-5 (measured vs fp64)
    for k in fx.nograd:
        assert p[k].grad is None, k


@pytest.mark.parametrize("name", ["spectral2d_pair", "spectral2d_pair_odd", "spectral2d_pair_nyquist"])
def test_spectral2d(name):
    _check(Fixture(name), lambda p, x: O.spectral_conv2d(x, p["weights1"], p["weights2"]))


def test_spectral2d_c64():
    _check(Fixture("spectral2d_c64"), lambda p, x: O.spectral_conv2d_c64(x, p["weights1"], p["weights2"]))


@pytest.mark.parametrize("name", ["spectral1d", "spectral1d_odd"])
def test_spectral1d(name):
    _check(Fixture(name), lambda p, x: O.spectral_conv1d(x, p["weights1"]))


@pytest.mark.parametrize("name", ["fno2d", "fno2d_rect"])
def test_fno2d(name):
    _check(Fixture(name), O.fno2d_forward)


@pytest.mark.parametrize("name", ["fno1d", "fno1d_banker"])
def test_fno1d(name):
    _check(Fixture(name), O.fno1d_forward)


@pytest.mark.parametrize("name,heads", [("niofp2d_fno_eval", ("fno_drift", "fno_diffusion")),
                                        ("niofp2d_nc_fno_eval", ("fno_Fx", "fno_Fy"))])
def test_niofp2d_fno_eval(name, heads):
    fx = Fixture(name)
    _check(fx, O.niofp2d_fno_forward, fx.t("meta.grid"), heads=heads)


def test_niofp2d_fno_train_bag_draw():
    fx = Fixture("niofp2d_fno_train")
    np.random.seed(int(fx.meta("np_seed")))
    idx = O.draw_bag(fx.t("x").shape[1], True)
    assert np.array_equal(idx, fx.meta("idx"))
    _check(fx, O.niofp2d_fno_forward, fx.t("meta.grid"), idx=idx)


def test_niofp1d_fno_train():
    fx = Fixture("niofp1d_fno_train")
    np.random.seed(int(fx.meta("np_seed")))
    idx = O.draw_bag(fx.t("x").shape[1], True)
    assert np.array_equal(idx, fx.meta("idx"))
    _check(fx, O.niofp1d_fno_forward, fx.t("meta.grid"), idx=idx)


def test_niofp1d_gpe_fno_eval():
    fx = Fixture("niofp1d_gpe_fno_eval")
    _check(fx, O.niofp1d_fno_forward, fx.t("meta.grid"), heads=("fno_V",))


def test_default_head_layer_from_seed():
    fx = Fixture("spectral2d_default_head")
    torch.manual_seed(int(fx.meta("weight_seed")))
    scale = 1.0 / (12 * 12)
    w1 = scale * torch.rand(12, 12, 32, 32, 2)
    w2 = scale * torch.rand(12, 12, 32, 32, 2)
    x = torch.randn(1, 12, 76, 76, generator=torch.Generator().manual_seed(int(fx.meta("x_seed"))))
    assert rel_err(O.spectral_conv2d(x, w1, w2), fx.t("y")) < TOL
    assert rel_err(dft64.spectral_conv2d(x.numpy(), w1.numpy(), w2.numpy()), fx.t("y")) < TOL


# ---- the fp64 pruned-DFT restatement against the same golden vectors --------
@pytest.mark.parametrize("name", ["spectral2d_pair", "spectral2d_pair_odd", "spectral2d_pair_nyquist", "spectral2d_c64"])
def test_dft64_2d(name):
    fx = Fixture(name)
    p = fx.params
    w1, w2 = p["weights1"].numpy(), p["weights2"].numpy()
    assert rel_err(dft64.spectral_conv2d(fx.t("x").numpy(), w1, w2), fx.t("y")) < TOL
    gx, gw1, gw2 = dft64.spectral_conv2d_grads(fx.t("x").numpy(), w1, w2, fx.t("gy").numpy())
    assert rel_err(gx, fx.t("gx")) < 5 * TOL
    g = fx.grads
    if g["weights1"].is_complex():
        gw1, gw2 = gw1[..., 0] + 1j * gw1[..., 1], gw2[..., 0] + 1j * gw2[..., 1]
    assert rel_err(gw1, g["weights1"]) < 5 * TOL
    assert rel_err(gw2, g["weights2"]) < 5 * TOL


@pytest.mark.parametrize("name", ["spectral1d", "spectral1d_odd"])
def test_dft64_1d(name):
    fx = Fixture(name)
    w = fx.params["weights1"].numpy()
    assert rel_err(dft64.spectral_conv1d(fx.t("x").numpy(), w), fx.t("y")) < TOL
    gx, gw = dft64.spectral_conv1d_grads(fx.t("x").numpy(), w, fx.t("gy").numpy())
    assert rel_err(gx, fx.t("gx")) < 5 * TOL
    assert rel_err(gw, fx.grads["weights1"]) < 5 * TOL


# ---- live comparison when the reference tree is mounted (build container only) --
@pytest.mark.skipif(not os.path.isdir("/root/reference/2d_FPE"), reason="reference tree not mounted")
def test_live_reference_default_shape_model():
    from tests.golden.make_golden import load_reference
    N2 = load_reference("2d_FPE", "NIOModules")
    torch.manual_seed(3)
    m = N2.NIOFP2D_FNO(2, 3, 100, 25, 3, 12, 32, 2).train()
    ax = np.linspace(-1, 1, 61, dtype=np.float32)
    grid = torch.tensor(np.stack(np.meshgrid(ax, ax, indexing="ij"), axis=2))
    x = torch.randn(1, 100, 61, 61, generator=torch.Generator().manual_seed(1))
    np.random.seed(2)
    want = m(x, grid)
    p = {k: v for k, v in m.state_dict().items() if not k.startswith("branch.")}
    np.random.seed(2)
    got = O.niofp2d_fno_forward(p, x, grid, idx=O.draw_bag(100, True))
    assert rel_err(got, want) < TOL


# ---------------------------------------------------------------------------------------------
# NIO (DeepONet branch + trunk) fixtures: weights rebuilt from the stored construction seed
# ---------------------------------------------------------------------------------------------
NIO_CASES = {
    "nio1d_gpe_train": ("1d_GPE", "NIOFP_schrodinger", ("fno_V",), True),
    "nio1d_fpe_eval": ("1d_FPE", "NIOFP", ("fno_drift", "fno_diffusion"), False),
    "nio2d_fpe_train": ("2d_FPE", "NIOFP2D", ("fno_drift", "fno_diffusion"), True),
}


def build_nio_from_fixture(name, device="cpu"):
    """The surface model of a NIO fixture with the reference's seeded initial weights."""
    from blindno_b200.surface import nio
    variant, cls, heads, training = NIO_CASES[name]
    fx = Fixture(name)
    torch.manual_seed(int(fx.meta("weight_seed")))
    args = tuple(int(v) for v in fx.meta("ctor"))
    extra = ("cpu",) if variant.startswith("1d") else ()
    model = nio.make_models(variant)[cls](*args, *extra)
    return fx, model.train(training), heads, training


@pytest.mark.parametrize("name", list(NIO_CASES))
def test_nio_oracle_matches_reference_golden(name):
    fx, model, heads, training = build_nio_from_fixture(name)
    variant = NIO_CASES[name][0]
    p = {k: v.detach().clone().requires_grad_((v.is_floating_point() or v.is_complex()) and "running" not in k)
         for k, v in model.state_dict().items() if not k.startswith("deeponet.branch.") and not k.startswith("deeponet.trunk.")}
    idx = fx.meta("idx") if training else None
    grid = fx.t("meta.grid")
    if variant.startswith("1d"):
        y = O.nio1d_forward(p, fx.t("x"), grid, heads=heads, idx=idx, training=training,
                            use_final_conv4=(variant != "1d_FPE"))
    else:
        y = O.nio2d_forward(p, fx.t("x"), grid, heads=heads, idx=idx, training=training)
    assert y.shape == fx.t("y").shape
    assert rel_err(y, fx.t("y")) < 2e-5
    y.backward(fx.t("gy"))
    for k, g in fx.grads.items():
        assert p[k].grad is not None, k
        assert rel_err(p[k].grad, g) < 5e-4, k
    norms = fx.group("gnorm.")
    biggest = max(v[0].item() for v in norms.values())
    for k, stats in norms.items():     # (conv biases before a train-mode BatchNorm have an exactly-zero gradient: noise only)
        got = p[k].grad.double()
        assert abs(got.norm().item() - stats[0].item()) <= 1e-3 * max(stats[0].item(), 1e-3 * biggest), k
    assert p["fc0.weight"].grad is None and p["fc0.bias"].grad is None


@pytest.mark.parametrize("fixture,heads", [("endmetric_2d_fpe", ("fno_drift", "fno_diffusion")),
                                           ("endmetric_2d_nc", ("fno_Fx", "fno_Fy"))])
def test_end_metric_relative_l2_of_the_oracle_matches_the_reference(fixture, heads):
    """The number the reference reports (drift / diffusion relative L2, eval_fno.py) on a fixed synthetic problem,
    at the initial weights and after 6 steps of the reference train loop: the oracle reproduces both."""
    from tests.helpers import end_metric
    fx = Fixture(fixture)
    p = {k: v.clone() for k, v in fx.params.items()}
    grid = fx.t("grid")
    fwd = lambda pp, x, g, **kw: O.niofp2d_fno_forward(pp, x, g, heads=heads, **kw)      # noqa: E731
    with torch.no_grad():
        m0 = end_metric(fx, lambda x: fwd(p, x, grid))
    assert np.abs(m0 - fx.arrays["metric0"]).max() < 5e-7          # unchanged to 6 decimals
    for v in p.values():
        v.requires_grad_(True)
    opt = torch.optim.Adam(O.trainable(p), lr=float(fx.meta("lr")))
    np.random.seed(int(fx.meta("np_seed")))
    losses = [O.train_step(p, opt, fwd, fx.t("x_train"), grid, fx.t("y_train")).item() for _ in range(6)]
    assert np.allclose(losses, fx.arrays["losses"], rtol=2e-5)
    with torch.no_grad():
        m1 = end_metric(fx, lambda x: fwd(p, x, grid))
    assert np.abs(m1 - fx.arrays["metric1"]).max() < 2e-5


def test_oracle_gradcheck_fp64():
    """torch.autograd.gradcheck of the fp64 restatement (finite differences vs its autograd), so that
    the gradients the CUDA path is compared with are themselves pinned by something other than autograd."""
    g = torch.Generator().manual_seed(0)
    x2 = torch.randn(2, 2, 6, 8, dtype=torch.float64, generator=g, requires_grad=True)
    w1 = (torch.rand(2, 3, 2, 3, 2, dtype=torch.float64, generator=g) / 4).requires_grad_(True)
    w2 = (torch.rand(2, 3, 2, 3, 2, dtype=torch.float64, generator=g) / 4).requires_grad_(True)
    assert torch.autograd.gradcheck(O.spectral_conv2d, (x2, w1, w2), eps=1e-6, atol=1e-7)
    x1 = torch.randn(2, 2, 12, dtype=torch.float64, generator=g, requires_grad=True)
    wr = torch.rand(2, 3, 4, 2, dtype=torch.float64, generator=g) / 4
    wc = torch.view_as_complex(wr).clone().requires_grad_(True)
    assert torch.autograd.gradcheck(O.spectral_conv1d, (x1, wc), eps=1e-6, atol=1e-7)

    fx = Fixture("fno2d")
    p = {k: (v.double() if not v.is_complex() else v.to(torch.complex128)) for k, v in fx.params.items()}
    names = ["fc0.weight", "conv_list.0.weight", "spectral_list.1.weights2", "fc1.bias", "fc2.weight"]
    x = fx.t("x")[:1, :7, :7].double().clone().requires_grad_(True)

    def f(xx, *ws):
        q = dict(p)
        q.update(dict(zip(names, ws)))
        return O.fno2d_forward(q, xx)

    leaves = [p[n].clone().requires_grad_(True) for n in names]
    assert torch.autograd.gradcheck(f, (x, *leaves), eps=1e-6, atol=1e-6, nondet_tol=0.0)
