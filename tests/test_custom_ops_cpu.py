"""CPU: the torch.library custom-op layer -- every op is registered under
``torch.ops.blindno_b200`` with the documented schema, the autograd formulas are wired with the right
structure (checked on the meta device: shapes only, no kernel runs), and CPU tensors are rejected."""
import numpy as np
import pytest
import torch

from blindno_b200 import ops
from blindno_b200.surface import fno, nio

NSO = torch.ops.blindno_b200
META = "meta"

PUBLIC = {
    "spectral_conv2d": "blindno_b200::spectral_conv2d(Tensor x, Tensor w1, Tensor w2, int m1, int m2, int prec=0) -> Tensor",
    "spectral_conv1d": "blindno_b200::spectral_conv1d(Tensor x, Tensor w, int m, bool halve_dc=True, int prec=0) -> Tensor",
    "fno_lift_pad": "blindno_b200::fno_lift_pad(Tensor x_cl, Tensor fc0_w, Tensor fc0_b, int ndim) -> Tensor",
    "fno_layer2d": "blindno_b200::fno_layer2d(Tensor z, Tensor w1, Tensor w2, Tensor conv_w, Tensor conv_b, bool gelu_in, int prec=0) -> Tensor",
    "fno_layer1d": "blindno_b200::fno_layer1d(Tensor z, Tensor w, Tensor conv_w, Tensor conv_b, bool gelu_in, int prec=0) -> Tensor",
    "fno_project": "blindno_b200::fno_project(Tensor z, Tensor fc1_w, Tensor fc1_b, Tensor fc2_w, Tensor fc2_b, int out_h, int out_w) -> Tensor",
    "bag_pool_lift": "blindno_b200::bag_pool_lift(Tensor s, Tensor grid, Tensor fc0_w, Tensor fc0_b) -> Tensor",
    "bag_project_pool_lift": None,
    "deeponet_pool_contract_lift": None,
    "fno_net": None,
    "adam_step_flat_": None,
}


def test_every_op_is_registered_with_its_schema():
    for name in ops.OP_NAMES:
        assert hasattr(NSO, name), name
    for name, schema in PUBLIC.items():
        op = getattr(NSO, name).default
        if schema is not None:
            assert str(op._schema) == schema
    # every differentiable op has a matching primitive pair
    for stem in ("spectral_conv", "fno_net", "fno_layer"):
        assert stem + "_forward" in ops.OP_NAMES and stem + "_backward" in ops.OP_NAMES
    assert str(NSO.adam_step_flat_.default._schema).count("!") == 3          # param, exp_avg, exp_avg_sq are mutated
    assert "Tensor(a!)? grad_sink" in str(NSO.fno_net_backward.default._schema)


def test_cpu_key_raises_for_every_primitive():
    x = torch.randn(1, 3, 8, 8)
    w = torch.rand(3, 3, 2, 2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        NSO.spectral_conv2d(x, w, w, 2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        NSO.fno_lift_pad(torch.randn(1, 8, 8, 3), torch.randn(4, 3), torch.randn(4), 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        NSO.bag_pool_lift(torch.randn(2, 3, 8), torch.randn(8, 1), torch.randn(4, 2), torch.randn(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        NSO.adam_step_flat_(*(torch.zeros(4) for _ in range(4)), 1e-3, 0.9, 0.999, 1e-8, 1, 1.0)


def test_mode_arguments_must_match_the_weights():
    x = torch.randn(1, 3, 8, 8, device=META)
    w = torch.rand(3, 3, 2, 2, 2, device=META)
    with pytest.raises(RuntimeError, match="modes"):
        NSO.spectral_conv2d(x, w, w, 3, 2)
    with pytest.raises(RuntimeError, match="DC"):
        NSO.spectral_conv1d(torch.randn(1, 3, 16, device=META), torch.rand(3, 3, 4, dtype=torch.cfloat, device=META), 4, False)


@pytest.mark.parametrize("ndim", [1, 2])
def test_spectral_autograd_structure_on_meta(ndim):
    if ndim == 2:
        x = torch.randn(2, 3, 8, 10, device=META, requires_grad=True)
        w1 = torch.rand(3, 5, 2, 3, 2, device=META, requires_grad=True)
        w2 = torch.rand(3, 5, 2, 3, 2, device=META, requires_grad=True)
        y = ops.spectral_conv(x, w1, w2)
        assert y.shape == (2, 5, 8, 10)
    else:
        x = torch.randn(2, 3, 16, device=META, requires_grad=True)
        w1 = torch.rand(3, 5, 4, dtype=torch.cfloat, device=META, requires_grad=True)
        w2 = None
        y = ops.spectral_conv(x, w1)
        assert y.shape == (2, 5, 16)
    y.sum().backward()
    assert x.grad.shape == x.shape and w1.grad.shape == w1.shape and w1.grad.dtype == w1.dtype
    if w2 is not None:
        assert w2.grad.shape == w2.shape
    # no spectrum is kept when nothing needs a gradient
    with torch.no_grad():
        _, xs = NSO.spectral_conv_forward(x, w1, w2, 0, False)
    assert xs.numel() == 0


def test_stage_chain_matches_whole_net_shapes_on_meta():
    m = fno.FNO2d(3, 4, 2, 3, 1).to(META)
    x = torch.randn(2, 9, 8, 3, device=META, requires_grad=True)
    whole = m(x)
    z = ops.fno_lift_pad(x, m.fc0.weight, m.fc0.bias, 2)
    assert z.shape == (2, 4, 9 + 2, 8 + 2)
    for k in range(2):
        s = m.spectral_list[k]
        z = ops.fno_layer(z, s.weights1, s.weights2, m.conv_list[k].weight, m.conv_list[k].bias, k > 0)
    # Q4: H is cropped by the W-derived pad and W by the H-derived pad
    out = ops.fno_project(z, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, 11 - 2, 10 - 2)
    assert out.shape == whole.shape == (2, 9, 8, 1)
    out.sum().backward()
    for name, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name
    assert x.grad.shape == x.shape

    m1 = fno.FNO1d(5, 6, 2, 2, 2).to(META)
    x1 = torch.randn(3, 20, 2, device=META)
    z = ops.fno_lift_pad(x1, m1.fc0.weight, m1.fc0.bias, 1)
    assert z.shape == (3, 6, 25)
    z = ops.fno_layer(z, m1.spectral_list[0].weights1, None, m1.conv_list[0].weight, m1.conv_list[0].bias, False)
    out = ops.fno_project(z, m1.fc1.weight, m1.fc1.bias, m1.fc2.weight, m1.fc2.bias, 1, 20)
    assert out.shape == m1(x1).shape == (3, 20, 2)
    out.sum().backward()
    assert m1.spectral_list[0].weights1.grad.dtype == torch.complex64


def test_nio_fno_model_autograd_structure_on_meta():
    model = nio.make_models("2d_FPE")["NIOFP2D_FNO"](2, 3, 100, 25, 2, 8, 3, 2).to(META).train()
    x = torch.randn(2, 60, 8, 8, device=META)
    grid = torch.randn(8, 8, 2, device=META)
    np.random.seed(0)
    pred = model(x, grid)
    assert pred.shape == (2, 8, 8, 2)
    pred.sum().backward()
    for name, p in model.named_parameters():
        if name.startswith("branch.") or name in ("fc0.weight", "fc0.bias"):
            assert p.grad is None, name              # unused branch; fc0 detached through .data (Q7)
        else:
            assert p.grad is not None and p.grad.shape == p.shape, name


def test_pooled_tails_detach_fc0_on_meta():
    grid = torch.randn(8, 8, 2, device=META)
    fc0_w = torch.randn(5, 3, device=META, requires_grad=True)
    fc0_b = torch.randn(5, device=META, requires_grad=True)
    fc1 = torch.nn.Linear(4, 128).to(META)
    fc2 = torch.nn.Linear(128, 1).to(META)
    z = torch.randn(6, 4, 10, 10, device=META, requires_grad=True)
    r = NSO.bag_project_pool_lift(z, 3, fc1.weight, fc1.bias, fc2.weight, fc2.bias, grid, fc0_w, fc0_b)
    assert r.shape == (2, 8, 8, 5)
    r.sum().backward()
    assert z.grad.shape == z.shape and fc1.weight.grad is not None and fc0_w.grad is None and fc0_b.grad is None

    w = torch.randn(2, 7, 25, device=META, requires_grad=True)
    basis = torch.randn(64, 25, device=META, requires_grad=True)
    b0 = torch.zeros((), device=META, requires_grad=True)
    r = NSO.deeponet_pool_contract_lift(w, basis, b0, grid, fc0_w, fc0_b)
    assert r.shape == (2, 8, 8, 5)
    r.sum().backward()
    assert w.grad.shape == w.shape and basis.grad.shape == basis.shape and b0.grad is not None and fc0_w.grad is None


def test_bag_attention_mean_on_meta():
    x = torch.randn(2, 9, 40, device=META, requires_grad=True)
    w = torch.ones(40, device=META, requires_grad=True)
    b = torch.zeros(40, device=META, requires_grad=True)
    y = ops.bag_attention_mean(x, w, b)
    assert y.shape == (2, 40)
    y.sum().backward()
    assert x.grad.shape == x.shape and w.grad.shape == w.shape and b.grad.shape == b.shape
    with pytest.raises(RuntimeError, match="128"):
        ops.bag_attention_mean(torch.randn(1, 200, 8, device=META), torch.ones(8, device=META), torch.zeros(8, device=META))


def test_heads_mse_on_meta():
    outs = [torch.randn(2, 8, 8, 1, device=META, requires_grad=True) for _ in range(2)]
    target = torch.randn(2, 8, 8, 2, device=META)
    loss = ops.heads_mse(outs, target)
    assert loss.shape == ()
    loss.backward()
    assert all(o.grad is not None and o.grad.shape == o.shape for o in outs)
    with pytest.raises(RuntimeError, match="target"):
        ops.heads_mse(outs, torch.randn(2, 8, 8, 3, device=META))


def test_grad_sink_suppresses_autograd_parameter_grads_on_meta():
    m = fno.FNO2d(3, 4, 1, 3, 1).to(META)
    params = m._params()
    sizes = [p.numel() * (2 if p.is_complex() else 1) for p in params]
    _, total = ops.slot_layout(sizes)
    sink = torch.zeros(total, device=META)
    out = ops.fno_apply(m._spec(), params, x_cl=torch.randn(2, 8, 8, 3, device=META), grad_sink=sink)
    out.sum().backward()
    assert all(p.grad is None for p in params)
