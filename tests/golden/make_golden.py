"""Generate golden input/output fixtures by running the UNMODIFIED reference modules.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It imports ``FNOModules`` / ``NIOModules`` read-only from the four reference
directories (with a 3-line ``timm`` stub, which the 2-D ``NIOModules`` imports
at module top but the hot path never uses), runs forward + backward in fp32 on
the CPU with seeded inputs and weights, and writes ``tests/golden/*.npz``.
Every fixture stores: inputs, the state_dict (minus the unused ``branch.*``
encoder of NIOFP2D_FNO, 9.9 M floats the path never touches), outputs, the
upstream gradient, and the gradient of every parameter autograd reached.
Nothing from the reference is copied; only tensors it produced are stored.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("BLINDNO_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
_MODNAMES = ("NIOModules", "FNOModules", "DeepONetModules", "Baselines", "debug_tools", "model", "utils")


def _timm_stub():
    if "timm" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def load_reference(subdir: str, module: str):
    """Import ``module`` from /root/reference/<subdir> with a clean module cache."""
    _timm_stub()
    for name in list(sys.modules):
        if name.split(".")[0] in _MODNAMES:
            del sys.modules[name]
    path = os.path.join(REF, subdir)
    sys.path.insert(0, path)
    try:
        return importlib.import_module(module)
    finally:
        sys.path.remove(path)


def _np(t):
    t = t.detach().cpu()
    return t.numpy().copy()


def _pack(prefix, named):
    return {f"{prefix}{k}": _np(v) for k, v in named}


def _save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays")


def _run(module, args, gy_seed):
    out = module(*args)
    g = torch.Generator().manual_seed(gy_seed)
    gy = torch.randn(out.shape, generator=g)
    out.backward(gy)
    return out, gy


def _module_case(name, module, x, extra_args=(), skip=("branch.",), **meta):
    x = x.clone().requires_grad_(True)
    out, gy = _run(module, (x, *extra_args), 99)
    params = [(k, v) for k, v in module.state_dict().items() if not k.startswith(skip)]
    grads = [(k, p.grad) for k, p in module.named_parameters() if p.grad is not None and not k.startswith(skip)]
    nograd = [k for k, p in module.named_parameters() if p.grad is None and not k.startswith(skip)]
    _save(name, x=_np(x), y=_np(out), gy=_np(gy), gx=_np(x.grad),
          nograd=np.array(nograd, dtype="U"),
          **{f"meta.{k}": np.asarray(v) for k, v in meta.items()},
          **_pack("p.", params), **_pack("g.", grads))


def main():
    torch.set_num_threads(1)
    g = torch.Generator()

    # ---- single spectral layers -------------------------------------------
    F2 = load_reference("2d_FPE", "FNOModules")
    torch.manual_seed(11)
    _module_case("spectral2d_pair", F2.SpectralConv2d(3, 4, 3, 4),
                 torch.randn(2, 3, 10, 12, generator=g.manual_seed(1)))
    torch.manual_seed(12)
    _module_case("spectral2d_pair_odd", F2.SpectralConv2d(2, 2, 4, 5),
                 torch.randn(3, 2, 9, 11, generator=g.manual_seed(2)))
    torch.manual_seed(13)
    _module_case("spectral2d_pair_nyquist", F2.SpectralConv2d(2, 3, 4, 5),
                 torch.randn(2, 2, 8, 8, generator=g.manual_seed(3)))

    F1 = load_reference("1d_FPE", "FNOModules")
    torch.manual_seed(14)
    _module_case("spectral1d", F1.SpectralConv1d(3, 3, 5),
                 torch.randn(2, 3, 20, generator=g.manual_seed(4)))
    torch.manual_seed(15)
    _module_case("spectral1d_odd", F1.SpectralConv1d(2, 4, 6),
                 torch.randn(3, 2, 15, generator=g.manual_seed(5)))
    torch.manual_seed(16)
    _module_case("spectral2d_c64", F1.SpectralConv2d(3, 3, 3, 4),
                 torch.randn(2, 3, 10, 12, generator=g.manual_seed(6)))

    # ---- FNO nets ------------------------------------------------------------
    torch.manual_seed(21)
    _module_case("fno2d", F2.FNO2d(4, 5, 3, 3, 1),
                 torch.randn(2, 13, 13, 3, generator=g.manual_seed(7)))
    torch.manual_seed(22)
    _module_case("fno2d_rect", F2.FNO2d(3, 4, 2, 2, 1),      # exercises the swapped crop (Q4)
                 torch.randn(2, 12, 17, 2, generator=g.manual_seed(8)))
    torch.manual_seed(23)
    _module_case("fno1d", F1.FNO1d(5, 6, 3, 2, 2),
                 torch.randn(3, 22, 2, generator=g.manual_seed(9)))       # round(5.5) -> 6
    torch.manual_seed(24)
    _module_case("fno1d_banker", F1.FNO1d(4, 3, 1, 1, 1),
                 torch.randn(2, 10, 1, generator=g.manual_seed(10)))      # round(2.5) -> 2

    # ---- whole NIO-FNO models -------------------------------------------------
    def grid2d(n):
        ax = np.linspace(-1, 1, n, dtype=np.float32)
        gx, gy = np.meshgrid(ax, ax, indexing="ij")
        return torch.tensor(np.stack([gx, gy], axis=2))

    N2 = load_reference("2d_FPE", "NIOModules")
    torch.manual_seed(31)
    m = N2.NIOFP2D_FNO(2, 3, 100, 25, 2, 6, 5, 2).eval()
    _module_case("niofp2d_fno_eval", m, torch.randn(2, 7, 20, 20, generator=g.manual_seed(11)),
                 extra_args=(grid2d(20),), grid=grid2d(20).numpy())
    m.train()
    np.random.seed(5)
    state = np.random.get_state()
    n_keep = np.random.randint(50, 53)
    idx = np.random.choice(53, n_keep)
    np.random.set_state(state)
    for prm in m.parameters():
        prm.grad = None
    _module_case("niofp2d_fno_train", m, torch.randn(1, 53, 20, 20, generator=g.manual_seed(12)),
                 extra_args=(grid2d(20),), grid=grid2d(20).numpy(), np_seed=5, idx=idx)

    NC = load_reference("2d_Non_conservative_FPE", "NIOModules")
    torch.manual_seed(32)
    m = NC.NIOFP2D_FNO(2, 3, 100, 25, 2, 5, 4, 2).eval()
    _module_case("niofp2d_nc_fno_eval", m, torch.randn(2, 5, 20, 20, generator=g.manual_seed(13)),
                 extra_args=(grid2d(20),), grid=grid2d(20).numpy())

    N1 = load_reference("1d_FPE", "NIOModules")
    torch.manual_seed(33)
    grid1 = torch.linspace(0, 1, 40).unsqueeze(-1)
    m = N1.NIOFP_FNO(2, 10, 7, 2, "cpu").train()
    np.random.seed(6)
    state = np.random.get_state()
    n_keep = np.random.randint(50, 54)
    idx = np.random.choice(54, n_keep)
    np.random.set_state(state)
    _module_case("niofp1d_fno_train", m, torch.randn(2, 54, 40, generator=g.manual_seed(14)),
                 extra_args=(grid1,), grid=grid1.numpy(), np_seed=6, idx=idx)

    NG = load_reference("1d_GPE", "NIOModules")
    torch.manual_seed(34)
    grid1 = torch.linspace(0, 1, 32).unsqueeze(-1)
    m = NG.NIOFP_FNO(3, 8, 9, 1, "cpu").eval()
    _module_case("niofp1d_gpe_fno_eval", m, torch.randn(2, 6, 32, generator=g.manual_seed(15)),
                 extra_args=(grid1,), grid=grid1.numpy())

    # ---- NIO models (DeepONet branch CNN + trunk FFN -> bag mean -> FNO heads) -------------
    # The encoders' default widths make the parameters large (1.5 M / 9.9 M floats), so these
    # fixtures store the construction SEED instead of the state_dict (tests rebuild the weights with
    # the surface classes, whose seeded init is identical: tests/test_surface_cpu.py), plus inputs,
    # outputs, and the gradients of everything but the conv stack, whose gradients are stored as norms.
    def _nio_case(name, module, x, grid, seed, **meta):
        out, gy = _run(module, (x, grid), 98)
        small = [(k, p.grad) for k, p in module.named_parameters()
                 if p.grad is not None and not k.startswith(("branch.conv", "branch.final_conv", "deeponet."))]
        big = [(k, torch.stack([p.grad.double().norm(), p.grad.double().sum()]))
               for k, p in module.named_parameters()
               if p.grad is not None and k.startswith(("branch.conv", "branch.final_conv"))]
        b0g = [("deeponet.b0", module.deeponet.b0.grad)]
        nograd = [k for k, p in module.named_parameters() if p.grad is None]
        _save(name, x=_np(x), y=_np(out), gy=_np(gy), nograd=np.array(nograd, dtype="U"),
              **{f"meta.{k}": np.asarray(v) for k, v in dict(meta, grid=grid.numpy(), weight_seed=seed).items()},
              **_pack("g.", small + b0g), **_pack("gnorm.", big))

    NG = load_reference("1d_GPE", "NIOModules")
    torch.manual_seed(51)
    m = NG.NIOFP_schrodinger(1, 3, 100, 25, 2, 8, 9, 1, "cpu").train()
    grid1 = torch.linspace(0, 1, 128).unsqueeze(-1)
    np.random.seed(7)
    state = np.random.get_state()
    n_keep = np.random.randint(50, 52)
    idx = np.random.choice(52, n_keep)
    np.random.set_state(state)
    _nio_case("nio1d_gpe_train", m, torch.randn(2, 52, 128, generator=g.manual_seed(21)), grid1, 51,
              np_seed=7, idx=idx, ctor=np.array([1, 3, 100, 25, 2, 8, 9, 1]))

    N1 = load_reference("1d_FPE", "NIOModules")
    torch.manual_seed(52)
    m = N1.NIOFP(1, 3, 100, 25, 2, 10, 7, 2, "cpu").eval()
    grid1 = torch.linspace(0, 1, 80).unsqueeze(-1)
    _nio_case("nio1d_fpe_eval", m, torch.randn(2, 5, 80, generator=g.manual_seed(22)), grid1, 52,
              ctor=np.array([1, 3, 100, 25, 2, 10, 7, 2]))

    N2 = load_reference("2d_FPE", "NIOModules")
    torch.manual_seed(53)
    m = N2.NIOFP2D(2, 3, 100, 25, 2, 6, 5, 2).train()
    np.random.seed(8)
    state = np.random.get_state()
    n_keep = np.random.randint(50, 51)
    idx = np.random.choice(51, n_keep)
    np.random.set_state(state)
    _nio_case("nio2d_fpe_train", m, torch.randn(1, 51, 61, 61, generator=g.manual_seed(23)), grid2d(61), 53,
              np_seed=8, idx=idx, ctor=np.array([2, 3, 100, 25, 2, 6, 5, 2]))

    # ---- one default-shape spectral layer (heads: C=12, 76x76, m=32), weights from a seed
    torch.manual_seed(41)
    layer = F2.SpectralConv2d(12, 12, 32, 32)
    x = torch.randn(1, 12, 76, 76, generator=g.manual_seed(16))
    _save("spectral2d_default_head", y=_np(layer(x)), **{"meta.weight_seed": np.asarray(41),
                                                         "meta.x_seed": np.asarray(16)})


def rel_l2(a, b, eps=1e-12):
    """The metric the reference's eval scripts report (2d_FPE/eval_fno.py:124-128): ||a-b||_2 / (||b||_2 + eps)."""
    return float(np.linalg.norm((a - b).ravel(), 2) / (np.linalg.norm(b.ravel(), 2) + eps))


def end_metric(variant="2d_FPE", name="endmetric_2d_fpe"):
    """The reference's reported end metric on a fixed synthetic problem: drift / diffusion relative L2 error of
    the de-normalised prediction (2d_FPE/eval_fno.py:72-97, :274-276) at the initial weights and after 6 steps
    of the reference train loop (2d_FPE/train_fno.py:115-143: Adam 5e-4, MSE, a fresh bag subsample per step)."""
    torch.set_num_threads(1)
    g = torch.Generator()
    N2 = load_reference(variant, "NIOModules")        # (2d_Non_conservative_FPE: the two channels are Fx, Fy)
    torch.manual_seed(61)
    ctor = (2, 3, 100, 25, 2, 6, 5, 2)
    model = N2.NIOFP2D_FNO(*ctor)
    n = 20
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    grid = torch.tensor(np.stack(np.meshgrid(ax, ax, indexing="ij"), axis=2))
    # smooth synthetic fields: the targets are fixed functions of the grid modulated per sample
    xx, yy = grid[..., 0], grid[..., 1]
    def fields(k):
        a = torch.randn(k, 1, 1, generator=g)
        b = torch.randn(k, 1, 1, generator=g)
        drift = a * torch.sin(3.0 * xx) * torch.cos(2.0 * yy) + 0.5 * b * xx * yy
        diff = 1.0 + 0.3 * b * torch.cos(2.0 * xx + yy) + 0.1 * a
        return torch.stack([drift, diff], dim=-1)                      # [k, n, n, 2]
    g.manual_seed(31)
    y_train, y_eval = fields(4), fields(3)
    x_train = torch.randn(4, 60, n, n, generator=g) + y_train[..., 0].unsqueeze(1)
    x_eval = torch.randn(3, 60, n, n, generator=g) + y_eval[..., 0].unsqueeze(1)
    stats = dict(drift_mean=np.float32(0.05), drift_std=np.float32(1.7), diff_mean=np.float32(1.0), diff_std=np.float32(0.4))

    def evaluate():
        model.eval()
        rows = []
        with torch.no_grad():
            for k in range(x_eval.shape[0]):
                pred = model(x_eval[k:k + 1], grid)
                truth = y_eval[k].numpy()
                row = []
                for c, nm in enumerate(("drift", "diff")):
                    p_raw = pred[0, ..., c].numpy() * stats[nm + "_std"] + stats[nm + "_mean"]
                    t_raw = truth[..., c] * stats[nm + "_std"] + stats[nm + "_mean"]
                    row.append(rel_l2(p_raw, t_raw))
                rows.append(row)
        return np.asarray(rows, dtype=np.float64)

    state0 = [(k, v.clone()) for k, v in model.state_dict().items() if not k.startswith("branch.")]
    metric0 = evaluate()
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    np.random.seed(9)
    losses = []
    for step in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(x_train, grid), y_train)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    metric1 = evaluate()
    _save(name, x_train=_np(x_train), y_train=_np(y_train), x_eval=_np(x_eval), y_eval=_np(y_eval),
          grid=_np(grid), metric0=metric0, metric1=metric1, losses=np.asarray(losses, dtype=np.float64),
          **{"meta.ctor": np.asarray(ctor), "meta.np_seed": np.asarray(9), "meta.lr": np.asarray(5e-4),
             "meta.variant": np.asarray(variant)},
          **{"stats." + k: np.asarray(v) for k, v in stats.items()}, **_pack("p.", state0))


def blindno_cases():
    """BlinDNO models: the reference's PermInvUNet_attn* classes, forward + backward.  The FNO
    heads are hard-wired to 32 (2-D) / 15 (1-D) modes, so the fixtures store the construction seed and keyword
    arguments instead of the state_dict, the full gradients of the small tensors and (norm, sum) of the others."""
    torch.set_num_threads(1)
    g = torch.Generator()

    def case(name, variant, cls, kwargs, x, seed, train, np_seed=None):
        mod = load_reference(variant, "NIOModules")
        torch.manual_seed(seed)
        model = getattr(mod, cls)(**kwargs)
        model.train(train)
        if np_seed is not None:
            np.random.seed(np_seed)
        out = model(x)
        gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(97))
        out.backward(gy)
        named = [(k, p) for k, p in model.named_parameters() if p.grad is not None]
        small = [(k, p.grad) for k, p in named if p.numel() <= 4096]
        real = lambda t: torch.view_as_real(t) if t.is_complex() else t
        big = [(k, torch.stack([real(p.grad).double().norm(), real(p.grad).double().sum()])) for k, p in named if p.numel() > 4096]
        nograd = [k for k, p in model.named_parameters() if p.grad is None]
        kw = {k: np.asarray(v) for k, v in kwargs.items() if k != "device"}
        _save(name, x=_np(x), y=_np(out), gy=_np(gy), nograd=np.array(nograd, dtype="U"),
              **{"meta.variant": np.asarray(variant), "meta.cls": np.asarray(cls), "meta.seed": np.asarray(seed),
                 "meta.train": np.asarray(train), "meta.np_seed": np.asarray(-1 if np_seed is None else np_seed)},
              **{"kw." + k: v for k, v in kw.items()}, **_pack("g.", small), **_pack("gnorm.", big))

    case("blindno2d_fpe_train", "2d_FPE", "PermInvUNet_attn", dict(base_ch=2, depth=2, input_size=(52, 52)),
         torch.randn(1, 53, 52, 52, generator=g.manual_seed(41)), 71, True, np_seed=11)
    case("blindno2d_nc_eval", "2d_Non_conservative_FPE", "PermInvUNet_attn", dict(base_ch=2, depth=2, input_size=(52, 53)),
         torch.randn(2, 6, 52, 53, generator=g.manual_seed(42)), 72, False)
    case("blindno1d_fpe_bag_train", "1d_FPE", "PermInvUNet_attn1D_bag", dict(base_ch=2, depth=3, input_size=45, device="cpu"),
         torch.randn(2, 54, 45, generator=g.manual_seed(43)), 73, True, np_seed=12)
    case("blindno1d_gpe_bag_eval", "1d_GPE", "PermInvUNet_attn1D_bag_GPE",
         dict(base_ch=2, depth=2, input_size=64, device="cpu", width=8, modes=9),
         torch.randn(2, 6, 64, generator=g.manual_seed(44)), 74, False)


if __name__ == "__main__":
    wanted = [a for a in sys.argv[1:] if a in ("end_metric", "blindno")]
    if not wanted:
        main()
    if not wanted or "end_metric" in wanted:
        end_metric()
        end_metric("2d_Non_conservative_FPE", "endmetric_2d_nc")
    if not wanted or "blindno" in wanted:
        blindno_cases()
