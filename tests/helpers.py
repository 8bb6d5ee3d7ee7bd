import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


class Fixture:
    """One tests/golden/*.npz file produced by tests/golden/make_golden.py."""

    def __init__(self, name):
        raw = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.arrays = {k: raw[k] for k in raw.files}

    def t(self, key):
        return torch.from_numpy(self.arrays[key])

    def group(self, prefix):
        return {k[len(prefix):]: torch.from_numpy(v) for k, v in self.arrays.items() if k.startswith(prefix)}

    @property
    def params(self):
        return self.group("p.")

    @property
    def grads(self):
        return self.group("g.")

    @property
    def nograd(self):
        return [str(s) for s in self.arrays.get("nograd", [])]

    def meta(self, key, default=None):
        return self.arrays.get("meta." + key, default)


def rel_err(a, b):
    """max|a-b| / max|b| -- the metric the tolerance in BASELINE.json is stated in."""
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    if a.is_complex() or b.is_complex():
        a, b = torch.view_as_real(a.to(torch.complex128)), torch.view_as_real(b.to(torch.complex128))
    a, b = a.double(), b.double()
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def assert_grads_close(got, want, tol, floor=1e-2):
    """Every gradient within ``tol`` of the reference, relative to max|ref| of that
    tensor, or -- for gradients that are tiny sums of cancelling terms -- relative to
    ``floor`` x the largest gradient entry of the whole model."""
    def _abs(t):
        t = torch.as_tensor(t).detach().cpu()
        return (torch.view_as_real(t.to(torch.complex128)) if t.is_complex() else t.double()).abs()
    gmax = max(_abs(w).max().item() for w in want.values())
    for k, w in want.items():
        assert got[k] is not None, f"missing gradient for {k}"
        g = torch.as_tensor(got[k]).detach().cpu()
        w = torch.as_tensor(w).detach().cpu()
        if w.is_complex() or g.is_complex():
            g, w = torch.view_as_real(g.to(torch.complex128)), torch.view_as_real(w.to(torch.complex128))
        assert g.shape == w.shape, (k, g.shape, w.shape)
        err = (g.double() - w.double()).abs().max().item()
        scale = max(w.double().abs().max().item(), floor * gmax)
        assert err <= tol * scale, f"{k}: err {err:.3e} > {tol:.1e} * {scale:.3e}"


def rel_l2(a, b, eps=1e-12):
    """The reference's reported end metric (2d_FPE/eval_fno.py:124-128): ||a-b||_2 / (||b||_2 + eps)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel(), 2) / (np.linalg.norm(b.ravel(), 2) + eps))


def end_metric(fx, predict):
    """Drift / diffusion relative L2 of the de-normalised predictions on the fixture's eval set
    (2d_FPE/eval_fno.py:72-97, :274-276).  ``predict(x[1,L,n,n]) -> [1,n,n,2]`` CPU tensor."""
    rows = []
    for k in range(fx.arrays["x_eval"].shape[0]):
        pred = predict(fx.t("x_eval")[k:k + 1]).detach().cpu().numpy()
        truth = fx.arrays["y_eval"][k]
        row = []
        for c, nm in enumerate(("drift", "diff")):
            std, mean = fx.arrays[f"stats.{nm}_std"], fx.arrays[f"stats.{nm}_mean"]
            row.append(rel_l2(pred[0, ..., c] * std + mean, truth[..., c] * std + mean))
        rows.append(row)
    return np.asarray(rows)
