"""fp64 pruned-DFT restatement of the spectral convolutions.  TEST INFRASTRUCTURE ONLY.

Library-independent second opinion for oracle/blindno_oracle.py: no FFT call,
only the sums of SURVEY.md section 3.5 evaluated in float64 with NumPy.  It follows
2d_FPE/FNOModules.py:156-178 (rfft2 -> two corner blocks -> irfft2) and
1d_FPE/FNOModules.py:47-59 (rfft -> DC*0.5 -> mix -> irfft) of the reference:

  kept rows  K = {0..m1-1} u {Hp-m1..Hp-1},  kept cols l = 0..m2-1
  X[b,i,k,l] = sum_{h,w} x[b,i,h,w] e^{-i 2pi (kh/Hp + lw/Wp)}
  Y[b,o,k,l] = sum_i X[b,i,k,l] Wc[i,o,k,l]
  y[b,o,h,w] = Re sum_{k,l} c_l Y[b,o,k,l] e^{+i 2pi (kh/Hp + lw/Wp)} / (Hp Wp),  c_0=1, c_l=2

(the C2R step drops Im of column l=0 after the H inverse; Nyquist is never
kept at the shapes used: m2 <= Wp/2 when Wp even).  The backward formulas are
the ones the CUDA kernels implement; tests check them against autograd of the
torch oracle.  Stage functions (wfwd, hfwd, mix, hinv, winv) mirror the CUDA
kernels one to one so every kernel can be tested alone.
"""
from __future__ import annotations

import numpy as np


def kept_rows(hp: int, m1: int) -> np.ndarray:
    return np.concatenate([np.arange(m1), np.arange(hp - m1, hp)])


def w_table(wp: int, m2: int) -> np.ndarray:
    """e^{-i 2pi l w / Wp}, shape [m2, Wp] complex128."""
    l = np.arange(m2)[:, None]
    w = np.arange(wp)[None, :]
    return np.exp(-2j * np.pi * ((l * w) % wp) / wp)


def h_table(hp: int, m1: int) -> np.ndarray:
    """e^{-i 2pi k h / Hp} for the kept rows, shape [2*m1, Hp]."""
    k = kept_rows(hp, m1)[:, None]
    h = np.arange(hp)[None, :]
    return np.exp(-2j * np.pi * ((k * h) % hp) / hp)


def col_weight(wp: int, m2: int) -> np.ndarray:
    """Hermitian doubling of the half spectrum: 1 for l=0 (and Nyquist), 2 otherwise."""
    c = np.full(m2, 2.0)
    c[0] = 1.0
    if wp % 2 == 0 and m2 > wp // 2:
        c[wp // 2] = 1.0
    return c


# ---- stages ---------------------------------------------------------------
def wfwd(x: np.ndarray, m2: int) -> np.ndarray:
    """[..., Wp] real -> [..., m2] complex."""
    return np.einsum("...w,lw->...l", x.astype(np.float64), w_table(x.shape[-1], m2))


def hfwd(x1: np.ndarray, m1: int) -> np.ndarray:
    """[..., Hp, m2] complex -> [..., 2*m1, m2] complex."""
    return np.einsum("kh,...hl->...kl", h_table(x1.shape[-2], m1), x1)


def mix(xs: np.ndarray, wc: np.ndarray) -> np.ndarray:
    """xs [B,Ci,K,m2] complex, wc [Ci,Co,K,m2] complex -> [B,Co,K,m2]."""
    return np.einsum("bikl,iokl->bokl", xs, wc)


def hinv(ys: np.ndarray, hp: int) -> np.ndarray:
    """[..., 2*m1, m2] -> [..., Hp, m2] (conjugate table, no scaling)."""
    m1 = ys.shape[-2] // 2
    return np.einsum("kh,...kl->...hl", np.conj(h_table(hp, m1)), ys)


def winv(z: np.ndarray, wp: int, col_scale: np.ndarray) -> np.ndarray:
    """[..., m2] complex -> [..., Wp] real: Re sum_l col_scale[l] z_l e^{+i 2pi l w/Wp}."""
    m2 = z.shape[-1]
    return np.real(np.einsum("...l,lw->...w", z * col_scale, np.conj(w_table(wp, m2))))


def weights_complex_2d(w1: np.ndarray, w2: np.ndarray) -> np.ndarray:
    """[Ci,Co,m1,m2,2] real pairs (or complex [Ci,Co,m1,m2]) x2 -> [Ci,Co,2*m1,m2] complex128."""
    def cplx(w):
        w = np.asarray(w)
        return w.astype(np.complex128) if np.iscomplexobj(w) else w[..., 0].astype(np.float64) + 1j * w[..., 1]
    return np.concatenate([cplx(w1), cplx(w2)], axis=2)


# ---- whole ops ------------------------------------------------------------
def spectral_conv2d(x, w1, w2):
    """x [B,Ci,Hp,Wp] -> [B,Co,Hp,Wp] float64."""
    x = np.asarray(x, dtype=np.float64)
    hp, wp = x.shape[-2:]
    wc = weights_complex_2d(w1, w2)
    m1, m2 = wc.shape[2] // 2, wc.shape[3]
    assert 2 * m1 <= hp and m2 <= wp // 2 + 1
    ys = mix(hfwd(wfwd(x, m2), m1), wc)
    return winv(hinv(ys, hp), wp, col_weight(wp, m2) / (hp * wp))


def spectral_conv2d_grads(x, w1, w2, gy):
    """Analytic backward (the formulas the kernels use).  Returns gx, gw1, gw2
    with gw* in the [Ci,Co,m1,m2,2] real-pair layout."""
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    hp, wp = x.shape[-2:]
    wc = weights_complex_2d(w1, w2)
    m1, m2 = wc.shape[2] // 2, wc.shape[3]
    xs = hfwd(wfwd(x, m2), m1)
    gys = hfwd(wfwd(gy, m2), m1) * (col_weight(wp, m2) / (hp * wp))
    gxs = np.einsum("bokl,iokl->bikl", gys, np.conj(wc))
    gws = np.einsum("bikl,bokl->iokl", np.conj(xs), gys)
    gx = winv(hinv(gxs, hp), wp, np.ones(m2))
    pair = lambda g: np.stack([g.real, g.imag], axis=-1)
    return gx, pair(gws[:, :, :m1]), pair(gws[:, :, m1:])


def spectral_conv1d(x, w):
    """x [B,Ci,Np], w [Ci,Co,m] complex -> [B,Co,Np] float64 (DC bin halved before the mix)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    w = np.asarray(w).astype(np.complex128)
    m = w.shape[-1]
    assert m <= n // 2 + 1
    dc = np.ones(m)
    dc[0] = 0.5
    ys = np.einsum("bil,iol->bol", wfwd(x, m) * dc, w)
    return winv(ys, n, col_weight(n, m) / n)


def spectral_conv1d_grads(x, w, gy):
    """Returns gx [B,Ci,Np] and gw [Ci,Co,m] complex (PyTorch convention: conj(X)*GY)."""
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    n = x.shape[-1]
    w = np.asarray(w).astype(np.complex128)
    m = w.shape[-1]
    dc = np.ones(m)
    dc[0] = 0.5
    xs = wfwd(x, m) * dc
    gys = wfwd(gy, m) * (col_weight(n, m) / n)
    gxs = np.einsum("bol,iol->bil", gys, np.conj(w)) * dc
    gw = np.einsum("bil,bol->iol", np.conj(xs), gys)
    return winv(gxs, n, np.ones(m)), gw
