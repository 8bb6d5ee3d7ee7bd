"""FNOModules.py surface: SpectralConv1d/2d, FNO1d, FNO2d.

Same class names, ctor signatures and state_dict layout as the reference
(1d_FPE/FNOModules.py:27-122, 2d_FPE/FNOModules.py:124-240); parameters are created in the same
order with the same initialisers, so a seeded construction yields the same weights.  ``forward``
hands the whole net to one autograd op over libblindno_b200.so instead of torch.fft / einsum /
conv / gelu calls.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class SpectralConv1d(nn.Module):
    """rfft -> DC bin * 0.5 -> first ``modes1`` bins mixed by a complex [Cin, Cout] matrix each -> irfft."""

    def __init__(self, in_channels, out_channels, modes1):
        super().__init__()
        self.in_channels, self.out_channels, self.modes1 = in_channels, out_channels, modes1
        self.scale = 1 / (in_channels * out_channels)
        self.weights1 = nn.Parameter(self.scale * torch.rand(in_channels, out_channels, modes1, dtype=torch.cfloat))

    def forward(self, x):
        return ops.spectral_conv(x, self.weights1)


class _SpectralConv2dBase(nn.Module):
    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1 / (in_channels * out_channels)

    def forward(self, x):
        return ops.spectral_conv(x, self.weights1, self.weights2)


class SpectralConv2d(_SpectralConv2dBase):
    """2-D dirs: weights are float32 [Cin, Cout, m1, m2, 2] (re, im) pairs (2d_FPE/FNOModules.py:138-139)."""

    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__(in_channels, out_channels, modes1, modes2)
        shape = (in_channels, out_channels, modes1, modes2, 2)
        self.weights1 = nn.Parameter(self.scale * torch.rand(*shape, dtype=torch.float32))
        self.weights2 = nn.Parameter(self.scale * torch.rand(*shape, dtype=torch.float32))


class SpectralConv2dC64(_SpectralConv2dBase):
    """1-D dirs' SpectralConv2d: complex64 [Cin, Cout, m1, m2] weights (1d_FPE/FNOModules.py:138-139)."""

    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__(in_channels, out_channels, modes1, modes2)
        shape = (in_channels, out_channels, modes1, modes2)
        self.weights1 = nn.Parameter(self.scale * torch.rand(*shape, dtype=torch.cfloat))
        self.weights2 = nn.Parameter(self.scale * torch.rand(*shape, dtype=torch.cfloat))


class _FnoBase(nn.Module):
    ndim = 0

    def _spec(self) -> ops.FnoSpec:
        m1 = self.modes1 if self.ndim == 2 else 0
        m2 = self.modes2 if self.ndim == 2 else self.modes
        return ops.FnoSpec(ndim=self.ndim, c_in=self.fc0.in_features, width=self.width,
                           c_out=self.fc2.out_features, n_layers=self.n_layers, modes1=m1, modes2=m2,
                           hidden=self.fc1.out_features, prec=getattr(self, "precision", ops.PREC_FP32))

    def _params(self):
        ps = [self.fc0.weight, self.fc0.bias]
        ps += [c.weight for c in self.conv_list] + [c.bias for c in self.conv_list]
        ps += [s.weights1 for s in self.spectral_list]
        if self.ndim == 2:
            ps += [s.weights2 for s in self.spectral_list]
        ps += [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias]
        return ps

    _grad_sink = None     # set by parallel.FlatTrainer: flat gradient buffer the backward kernels write into

    def forward(self, x):
        return ops.fno_apply(self._spec(), self._params(), x_cl=x, grad_sink=self._grad_sink)

    def encode_bags(self, bags, grid, idx=None, pool=None):
        """The per-snapshot NIO-FNO encoder: every snapshot of every bag, concatenated with the grid, through
        this net; with ``pool=(fc0.weight, fc0.bias)`` followed by the bag mean and the detached lift."""
        return ops.fno_apply(self._spec(), self._params(), bags=bags, grid=grid, idx=idx, pool=pool,
                             grad_sink=self._grad_sink)


class FNO1d(_FnoBase):
    ndim = 1

    def __init__(self, modes, width, n_layers, input_dim, output_dim, device="cpu"):
        super().__init__()
        self.modes, self.width, self.n_layers = modes, width, n_layers
        self.fc0 = nn.Linear(input_dim, width)
        self.conv_list = nn.ModuleList([nn.Conv1d(width, width, 1) for _ in range(n_layers)])
        self.spectral_list = nn.ModuleList([SpectralConv1d(width, width, modes) for _ in range(n_layers)])
        self.padding_frac = 1 / 4
        self.fc1 = nn.Linear(width, 128)
        self.fc2 = nn.Linear(128, output_dim)
        self.to(device)


class _Fno2dBase(_FnoBase):
    ndim = 2
    _spectral_cls = SpectralConv2d

    def _build(self, modes, width, n_layers, input_dim):
        self.modes1 = self.modes2 = modes
        self.width, self.n_layers = width, n_layers
        self.padding_frac = 1 / 4
        self.fc0 = nn.Linear(input_dim, width)
        self.conv_list = nn.ModuleList([nn.Conv2d(width, width, 1) for _ in range(n_layers)])
        self.spectral_list = nn.ModuleList([self._spectral_cls(width, width, modes, modes) for _ in range(n_layers)])
        self.fc1 = nn.Linear(width, 128)
        self.fc2 = nn.Linear(128, 1)     # Q3: the reference ignores output_dim here


class FNO2d(_Fno2dBase):
    """2d_FPE / 2d_Non_conservative_FPE signature (no device argument), real-pair spectral weights."""

    def __init__(self, modes, width, n_layers, input_dim, output_dim):
        super().__init__()
        self._build(modes, width, n_layers, input_dim)


class FNO2dC64(_Fno2dBase):
    """1d_FPE / 1d_GPE signature (extra ``device``), complex64 spectral weights."""
    _spectral_cls = SpectralConv2dC64

    def __init__(self, modes, width, n_layers, input_dim, output_dim, device="cpu"):
        super().__init__()
        self._build(modes, width, n_layers, input_dim)
        self.to(device)
