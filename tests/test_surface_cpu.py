"""CPU: the drop-in module surface -- names, shapes, dtypes, registration order, RNG consumption,
checkpoint loading -- against the golden fixtures (and the live reference when it is mounted)."""
import os

import numpy as np
import pytest
import torch

from blindno_b200.surface import fno, nio
from tests.helpers import Fixture


def _build(name):
    if name == "niofp2d_fno_eval" or name == "niofp2d_fno_train":
        return nio.make_models("2d_FPE")["NIOFP2D_FNO"](2, 3, 100, 25, 2, 6, 5, 2)
    if name == "niofp2d_nc_fno_eval":
        return nio.make_models("2d_Non_conservative_FPE")["NIOFP2D_FNO"](2, 3, 100, 25, 2, 5, 4, 2)
    if name == "niofp1d_fno_train":
        return nio.make_models("1d_FPE")["NIOFP_FNO"](2, 10, 7, 2, "cpu")
    if name == "niofp1d_gpe_fno_eval":
        return nio.make_models("1d_GPE")["NIOFP_FNO"](3, 8, 9, 1, "cpu")
    raise KeyError(name)


@pytest.mark.parametrize("name", ["niofp2d_fno_eval", "niofp2d_nc_fno_eval", "niofp1d_fno_train", "niofp1d_gpe_fno_eval"])
def test_state_dict_matches_reference_checkpoint_layout(name):
    fx = Fixture(name)
    model = _build(name)
    mine = {k: v for k, v in model.state_dict().items() if not k.startswith("branch.")}
    ref = fx.params
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert mine[k].shape == ref[k].shape and mine[k].dtype == ref[k].dtype, k
    missing, unexpected = model.load_state_dict(ref, strict=False)
    assert not unexpected and all(k.startswith("branch.") for k in missing)
    # DDP checkpoints carry a "module." prefix that the eval scripts strip (2d_FPE/eval_fno.py:104-122)
    stripped = {k[len("module."):]: v for k, v in {"module." + k: v for k, v in ref.items()}.items()}
    model.load_state_dict(stripped, strict=False)


def test_fno_layouts():
    m = fno.FNO2d(32, 12, 3, 12, 1)
    sd = m.state_dict()
    assert sd["spectral_list.0.weights1"].shape == (12, 12, 32, 32, 2) and sd["spectral_list.0.weights1"].dtype == torch.float32
    assert sd["conv_list.2.weight"].shape == (12, 12, 1, 1) and sd["fc2.weight"].shape == (1, 128)
    assert fno.FNO2d(4, 4, 1, 3, 7).fc2.out_features == 1          # Q3: output_dim ignored in 2-D
    m1 = fno.FNO1d(15, 30, 3, 30, 2)
    assert m1.state_dict()["spectral_list.1.weights1"].dtype == torch.complex64
    assert m1.state_dict()["spectral_list.1.weights1"].shape == (30, 30, 15) and m1.fc2.out_features == 2
    c = fno.FNO2dC64(3, 4, 1, 2, 1, device="cpu")
    assert c.state_dict()["spectral_list.0.weights2"].dtype == torch.complex64


def test_bag_draw_consumes_numpy_stream_like_the_reference():
    for name in ("niofp2d_fno_train", "niofp1d_fno_train"):
        fx = Fixture(name)
        np.random.seed(int(fx.meta("np_seed")))
        idx = nio.draw_bag(fx.t("x").shape[1], True)
        assert np.array_equal(idx, fx.meta("idx"))
        after = np.random.randint(0, 1 << 30)
        np.random.seed(int(fx.meta("np_seed")))
        n = np.random.randint(50, fx.t("x").shape[1]); np.random.choice(fx.t("x").shape[1], n)
        assert after == np.random.randint(0, 1 << 30)
    state = np.random.get_state()[1].copy()
    assert nio.draw_bag(100, False) is None
    assert np.array_equal(state, np.random.get_state()[1])       # eval draws nothing


def test_cpu_tensors_are_rejected_loudly():
    m = fno.FNO2d(3, 4, 2, 3, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 8, 8, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        fno.SpectralConv1d(2, 2, 3)(torch.randn(1, 2, 16))


@pytest.mark.skipif(not os.path.isdir("/root/reference/2d_FPE"), reason="reference tree not mounted")
@pytest.mark.parametrize("variant,cls,args", [
    ("2d_FPE", "NIOFP2D_FNO", (2, 3, 100, 25, 3, 12, 32, 2)),
    ("2d_Non_conservative_FPE", "NIOFP2D_FNO", (2, 3, 100, 25, 3, 12, 32, 2)),
    ("1d_FPE", "NIOFP_FNO", (3, 30, 15, 2, "cpu")),
    ("1d_GPE", "NIOFP_FNO", (3, 20, 40, 1, "cpu")),
    ("1d_GPE", "NIOFP_schrodinger", (1, 3, 100, 25, 3, 20, 40, 1, "cpu")),
    ("1d_FPE", "NIOFP", (1, 3, 100, 25, 3, 30, 15, 2, "cpu")),
    ("2d_FPE", "NIOFP2D", (2, 3, 100, 25, 3, 12, 32, 2)),
])
def test_live_state_dict_and_seeded_init_parity(variant, cls, args):
    from tests.golden.make_golden import load_reference
    R = load_reference(variant, "NIOModules")
    torch.manual_seed(7)
    ref = getattr(R, cls)(*args)
    torch.manual_seed(7)
    mine = nio.make_models(variant)[cls](*args)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]


# ---------------------------------------------------------------------------------------------
# drop-in module files: the imports the unchanged reference scripts perform
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant,imports", [
    ("2d_FPE", {"NIOModules": ["NIOFP2D", "NIOFP2D_FNO", "NIOFP2D_FNO_attn"]}),                 # 2d_FPE/train_fno.py:8
    ("2d_Non_conservative_FPE", {"NIOModules": ["NIOFP2D", "NIOFP2D_FNO", "PermInvUNet_attn"]}),
    ("1d_FPE", {"NIOModules": ["NIOFP", "NIOFP_FNO"], "FNOModules": ["FNO1d", "FNO2d", "FNO3d"]}),  # 1d_FPE/train_fno.py:7
    ("1d_GPE", {"NIOModules": ["NIOFP_schrodinger"], "Baselines": ["Encoder", "Encoder2D", "Encoder3D"],
                "DeepONetModules": ["FeedForwardNN", "DeepOnetNoBiasOrg", "FFN"]}),               # 1d_GPE/train_nio_GPE.py:7
])
def test_dropin_modules_resolve_the_scripts_imports(variant, imports):
    import importlib
    import subprocess
    import sys
    from blindno_b200 import dropin
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim_dir = os.path.join(os.path.dirname(dropin.__file__), variant)
    # (1) through PYTHONPATH, in a clean interpreter, exactly as a script's `from NIOModules import ...`
    lines = [f"from {m} import {', '.join(names)}" for m, names in imports.items()]
    lines.append("import NIOModules; print(sorted(n for n in dir(NIOModules) if n.startswith('NIOFP')))")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([shim_dir, root]))
    res = subprocess.run([sys.executable, "-c", "\n".join(lines)], capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    # (2) through install()
    saved = {k: sys.modules.get(k) for k in ("NIOModules", "FNOModules", "DeepONetModules", "Baselines", "debug_tools")}
    try:
        dropin.install(variant)
        for m, names in imports.items():
            mod = importlib.import_module(m)
            for n in names:
                assert hasattr(mod, n), (m, n)
        accelerated = nio.make_models(variant)
        import NIOModules
        for n in [n for n in imports["NIOModules"] if n in accelerated]:
            assert getattr(NIOModules, n).__name__ == n
        with pytest.raises(NotImplementedError):
            dropin.exports(variant, "NIOModules")["PermInvUNet"]()          # plain U-Net: out of scope, fails loudly
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ---------------------------------------------------------------------------------------------
# BlinDNO family: same state_dict, same seeded initial weights, same encoder
# ---------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.isdir("/root/reference/2d_FPE"), reason="reference tree not mounted")
@pytest.mark.parametrize("variant,cls,kwargs,xshape", [
    ("2d_FPE", "PermInvUNet_attn", dict(base_ch=2, depth=2, input_size=(21, 18)), (2, 55, 21, 18)),
    ("2d_Non_conservative_FPE", "PermInvUNet_attn", dict(base_ch=2, depth=3, input_size=(21, 18)), (2, 55, 21, 18)),
    ("1d_FPE", "PermInvUNet_attn1D_bag", dict(base_ch=2, depth=3, input_size=45, device="cpu"), (2, 55, 45)),
    ("1d_FPE", "PermInvUNet_attn1D", dict(base_ch=1, depth=2, input_size=40, device="cpu"), (2, 7, 40)),
    ("1d_GPE", "PermInvUNet_attn1D_bag", dict(base_ch=2, depth=2, input_size=64, device="cpu"), (2, 52, 64)),
    ("1d_GPE", "PermInvUNet_attn1D_bag_GPE", dict(base_ch=2, depth=2, input_size=64, device="cpu", width=8, modes=9), (2, 52, 64)),
])
def test_blindno_surface_matches_live_reference(variant, cls, kwargs, xshape):
    """Seeded construction gives bit-identical state_dicts (names, order, shapes, values); with the FNO heads
    replaced by the same stub in both, the U-Net encoder + bag attention agree and the bag draw consumes the
    NumPy stream identically (the heads themselves are CUDA-only: tests/test_gpu_parity.py)."""
    from blindno_b200.surface import blindno
    from tests.golden.make_golden import load_reference
    ref_mod = load_reference(variant, "NIOModules")
    torch.manual_seed(5)
    ref = getattr(ref_mod, cls)(**kwargs)
    torch.manual_seed(5)
    ours = blindno.make_blindno_models(variant)[cls](**kwargs)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k

    class Stub(torch.nn.Module):
        def forward(self, x):
            return x[..., :1] * 2.0

    for m in (ref, ours):
        for name in [n for n, _ in m.named_children() if n.startswith("fno_")]:
            setattr(m, name, Stub())
    x = torch.randn(*xshape, generator=torch.Generator().manual_seed(1))
    for training in (False, True):
        ref.train(training), ours.train(training)
        np.random.seed(3)
        want = ref(x)
        after_ref = np.random.randint(0, 1 << 30)
        np.random.seed(3)
        got = ours(x)
        assert np.random.randint(0, 1 << 30) == after_ref
        assert got.shape == want.shape
        assert (got - want).abs().max().item() <= 2e-5 * max(want.abs().max().item(), 1.0)


def test_dropin_exports_blindno_models():
    from blindno_b200 import dropin
    for variant, names in (("2d_FPE", ["PermInvUNet_attn"]), ("2d_Non_conservative_FPE", ["PermInvUNet_attn"]),
                           ("1d_FPE", ["PermInvUNet_attn1D", "PermInvUNet_attn1D_bag"]),
                           ("1d_GPE", ["PermInvUNet_attn1D_bag", "PermInvUNet_attn1D_bag_GPE"])):
        ex = dropin.exports(variant, "NIOModules")
        for n in names:
            assert isinstance(ex[n], type) and issubclass(ex[n], torch.nn.Module), (variant, n)
