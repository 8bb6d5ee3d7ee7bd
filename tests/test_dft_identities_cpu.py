"""CPU: the DFT identities the round-2 kernels rest on, in NumPy fp64 against numpy.fft.

* csrc/spectral.cu core2d / core2d_stream: the kept rows of the H transform (k = 0..m1-1 and k = -m1..-1 of
  SpectralConv2d, 2d_FPE/FNOModules.py:156-178) come in conjugate pairs that share cos / sin, so both H transforms are
  computed per frequency f = 0..m1.
* csrc/spectral.cu wfwd_fold: x[w] and x[wp - w] share cos and differ in the sign of sin.
"""
import numpy as np
import pytest


@pytest.mark.parametrize("hp,m1", [(76, 12), (76, 32), (100, 50), (37, 5), (8, 4)])
def test_h_transform_by_conjugate_row_pairs(hp, m1):
    rng = np.random.default_rng(hp * 100 + m1)
    x = rng.standard_normal(hp) + 1j * rng.standard_normal(hp)          # one column of the W-transformed image
    K = 2 * m1
    full = np.fft.fft(x)
    kept = np.concatenate([full[:m1], full[hp - m1:]])                  # rows k = 0..m1-1, then -m1..-1
    h = np.arange(hp)
    got = np.zeros(K, dtype=complex)
    for f in range(m1 + 1):
        c, s = np.cos(2 * np.pi * f * h / hp), np.sin(2 * np.pi * f * h / hp)
        A, B = (x * c).sum(), (x * s).sum()
        if f < m1:
            got[f] = A - 1j * B
        if f > 0:
            got[K - f] = A + 1j * B
    assert np.allclose(got, kept, rtol=1e-12, atol=1e-12)

    # inverse: z_h = sum_k Y[k] e^{+i phi_kh} over the kept rows = sum_f P_f cos + Q_f sin
    y = rng.standard_normal(K) + 1j * rng.standard_normal(K)
    freq = np.concatenate([np.arange(m1), np.arange(-m1, 0)])
    want = (y[:, None] * np.exp(2j * np.pi * freq[:, None] * h[None, :] / hp)).sum(axis=0)
    z = np.zeros(hp, dtype=complex)
    for f in range(m1 + 1):
        y1 = y[f] if f < m1 else 0.0
        y2 = y[K - f] if f > 0 else 0.0
        P, Q = y1 + y2, 1j * (y1 - y2)
        z += P * np.cos(2 * np.pi * f * h / hp) + Q * np.sin(2 * np.pi * f * h / hp)
    assert np.allclose(z, want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("wp,m2", [(76, 12), (76, 32), (100, 40), (160, 64), (36, 7), (7, 3)])
def test_w_forward_folded(wp, m2):
    rng = np.random.default_rng(wp * 100 + m2)
    x = rng.standard_normal(wp)
    want = np.fft.rfft(x)[:m2]
    nh = wp // 2 + 1
    e, o = np.zeros(nh), np.zeros(nh)
    for j in range(nh):
        jm = wp - j
        if j > 0 and jm > j:
            e[j], o[j] = x[j] + x[jm], x[j] - x[jm]
        else:                       # j = 0, and j = wp / 2 when wp is even: no partner
            e[j], o[j] = x[j], 0.0
    l, w = np.arange(m2)[:, None], np.arange(nh)[None, :]
    theta = 2 * np.pi * l * w / wp
    got = (e[None, :] * np.cos(theta)).sum(axis=1) - 1j * (o[None, :] * np.sin(theta)).sum(axis=1)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
