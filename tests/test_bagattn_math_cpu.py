"""CPU: the L x L algebra behind csrc/bagattn.cu, restated in torch fp64 and checked against autograd of the reference
formulation (TemporalSelfAttention.forward + .mean(dim=1), 2d_FPE/NIOModules.py:1063-1083, :1153-1170).

The CUDA kernels never form an [L, D] intermediate: between two passes over the tokens X everything is expressed through
the centered Gram matrix, the attention matrix and a few L-vectors.  This file pins those identities (forward and
backward) independently of the GPU, so a change of the kernels' formulas has a CPU oracle to answer to."""
import math

import pytest
import torch
import torch.nn.functional as F


def reference(x, gamma, beta, eps):
    d = x.shape[-1]
    attn = torch.softmax(torch.matmul(x, x.transpose(1, 2)) / math.sqrt(d), dim=-1)
    return F.layer_norm(torch.matmul(attn, x) + x, (d,), gamma, beta, eps).mean(dim=1)


def forward_algebra(x, gamma, beta, eps):
    """One bag [L, D] -> (out [D], saved)."""
    L, D = x.shape
    m = x.mean(dim=1)
    xc = x - m[:, None]
    gc = xc @ xc.T                                           # centered Gram matrix
    s = (gc + D * torch.outer(m, m)) / math.sqrt(D)          # = X X^T / sqrt(D)
    a = torch.softmax(s, dim=-1)
    mm = a + torch.eye(L, dtype=x.dtype)
    mu = mm @ m                                              # row means of O = M X
    var = torch.einsum("lk,kj,lj->l", mm, gc, mm) / D        # row variances of O: (M Gc M^T)_ll / D
    r = torch.rsqrt(var + eps)
    v = mm.T @ r
    c = (r * mu).sum()
    out = gamma / L * (v @ x - c) + beta
    return out, dict(m=m, gc=gc, a=a, mm=mm, mu=mu, r=r, v=v, c=c)


def backward_algebra(x, gamma, g, sv):
    """d loss / d x, d gamma, d beta of one bag from g = d loss / d out."""
    L, D = x.shape
    mm, gc, a, m, mu, r, v, c = sv["mm"], sv["gc"], sv["a"], sv["m"], sv["mu"], sv["r"], sv["v"], sv["c"]
    h = g * gamma / L
    hs = h.sum()
    hbar = hs / D
    wp = x @ h                                               # w' = X h
    al = r ** 3 / D * (mm @ wp - mu * hs)                    # a_l = r_l^2 p_l
    w = wp - hbar * D * m
    t = mm @ gc
    da = torch.outer(r, w) - al[:, None] * t
    ds = a * (da - (a * da).sum(dim=1, keepdim=True))
    k = -(mm.T * al[None, :]) @ mm + (ds + ds.T) / math.sqrt(D)
    z = mm.T @ (al * mu)
    dx = k @ x + torch.outer(v, h - hbar) + z[:, None]
    dgamma = g * (v @ x - c) / L
    return dx, dgamma, g.clone()


@pytest.mark.parametrize("L,D,shift", [(7, 37, 0.0), (12, 50, 1.5), (33, 20, -0.7), (5, 3, 0.0)])
def test_forward_and_backward_identities(L, D, shift):
    g = torch.Generator().manual_seed(L * 100 + D)
    x = (torch.randn(2, L, D, generator=g, dtype=torch.float64) * 0.8 + shift).requires_grad_(True)
    gamma = (1.0 + 0.3 * torch.randn(D, generator=g, dtype=torch.float64)).requires_grad_(True)
    beta = (0.2 * torch.randn(D, generator=g, dtype=torch.float64)).requires_grad_(True)
    gy = torch.randn(2, D, generator=g, dtype=torch.float64)
    eps = 1e-5
    want = reference(x, gamma, beta, eps)
    want.backward(gy)
    dgamma = torch.zeros(D, dtype=torch.float64)
    dbeta = torch.zeros(D, dtype=torch.float64)
    for b in range(2):
        out, sv = forward_algebra(x[b].detach(), gamma.detach(), beta.detach(), eps)
        assert torch.allclose(out, want[b].detach(), rtol=1e-10, atol=1e-12)
        dx, dg, db = backward_algebra(x[b].detach(), gamma.detach(), gy[b], sv)
        assert torch.allclose(dx, x.grad[b], rtol=1e-8, atol=1e-11)
        dgamma += dg
        dbeta += db
    assert torch.allclose(dgamma, gamma.grad, rtol=1e-9, atol=1e-12)
    assert torch.allclose(dbeta, beta.grad, rtol=1e-9, atol=1e-12)


def test_centered_gram_keeps_the_variance_a_sum_of_squares():
    """With token means far from zero the uncentered form E[O^2] - mu^2 loses digits in fp32; the centered Gram form
    does not (it is what the kernels use)."""
    g = torch.Generator().manual_seed(3)
    L, D = 9, 64
    x = (torch.randn(L, D, generator=g, dtype=torch.float64) * 0.1 + 50.0)
    _, sv = forward_algebra(x, torch.ones(D, dtype=torch.float64), torch.zeros(D, dtype=torch.float64), 1e-5)
    o = sv["mm"] @ x
    true_var = o.var(dim=1, unbiased=False)
    x32 = x.float()
    m32 = x32.mean(dim=1)
    xc32 = x32 - m32[:, None]
    mm32 = sv["mm"].float()
    centered = torch.einsum("lk,kj,lj->l", mm32, xc32 @ xc32.T, mm32) / D
    o32 = mm32 @ x32
    naive = (o32 * o32).mean(dim=1) - o32.mean(dim=1) ** 2
    err_centered = ((centered.double() - true_var).abs() / true_var).max().item()
    err_naive = ((naive.double() - true_var).abs() / true_var).max().item()
    assert err_centered < 1e-4 and err_naive > 10 * err_centered
