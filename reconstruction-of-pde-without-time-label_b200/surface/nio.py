"""NIOModules.py surface: the permutation-invariant bag models on the named hot path.

  NIO-FNO  NIOFP_FNO (1d_FPE/NIOModules.py:87-155, 1d_GPE/NIOModules.py:228-289)
           NIOFP2D_FNO (2d_FPE/NIOModules.py:508-581, 2d_Non_conservative_FPE/NIOModules.py:503-577)
  NIO      NIOFP (1d_FPE/NIOModules.py:15-84), NIOFP_schrodinger (1d_GPE/NIOModules.py:160-223),
           NIOFP2D (2d_FPE/NIOModules.py:14-83)

Same ctor signatures, attribute / state_dict names and registration order as the reference
(including the Encoder2D branch NIOFP2D_FNO builds and never calls, Q8).  Bag subsampling draws
from the global NumPy stream exactly like the reference (randint, then choice, Q6); ``fc0`` is
used through ``.data`` so it never receives a gradient (Q7).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops  # noqa: F401  (registers torch.ops.blindno_b200)
from .baselines import make_encoders
from .deeponet import FFN, DeepOnetNoBiasOrg
from .fno import FNO1d, FNO2d


def draw_bag(n_snapshots: int, training: bool):
    """Training: keep L ~ randint(50, L0) snapshots drawn with replacement (same idx for the whole batch)."""
    if not training:
        return None
    n_keep = np.random.randint(50, n_snapshots)
    return np.random.choice(n_snapshots, n_keep)


def _idx_tensor(idx, device):
    if idx is None:
        return None
    return torch.as_tensor(np.ascontiguousarray(idx), dtype=torch.int32).to(device, non_blocking=True)


_SIDE = {}


def _side_streams(device, n):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    pool = _SIDE.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


class _BagModel(nn.Module):
    head_names: tuple = ()
    accepts_idx = False     # forward(x, grid, idx=<int32 device tensor>): the caller drew the bag (CUDA-graph replay)
    _heads_as_list = False  # the trainer's fused loss (ops.heads_mse) reads the head outputs without the torch.cat

    def _heads(self, lifted):
        """The output FNO heads read the same lifted bag mean and are independent of each other; with only
        B images each they cannot fill the GPU alone, so every head after the first runs on its own side
        stream (fork / join by events -- capturable in a CUDA graph; autograd replays each head's backward
        on the stream its forward ran on)."""
        names = self.head_names
        if len(names) == 1 or not lifted.is_cuda:
            outs = [getattr(self, name)(lifted) for name in names]
            if self._heads_as_list:
                return outs
            return outs[0] if len(outs) == 1 else torch.cat(outs, dim=-1)
        main = torch.cuda.current_stream(lifted.device)
        side = _side_streams(lifted.device, len(names) - 1)
        outs = [None] * len(names)
        for k, name in enumerate(names[1:], start=1):
            st = side[k - 1]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                outs[k] = getattr(self, name)(lifted)
        outs[0] = getattr(self, names[0])(lifted)
        capturing = torch.cuda.is_current_stream_capturing()
        for k in range(1, len(names)):
            main.wait_stream(side[k - 1])
            if not capturing:          # (a graph's private pool keeps its tensors alive by itself)
                lifted.record_stream(side[k - 1])
                outs[k].record_stream(main)
        return outs if self._heads_as_list else torch.cat(outs, dim=-1)


class _NioFnoMixin(_BagModel):
    """FNO_input on every snapshot -> mean over the bag folded into the detached fc0 -> FNO heads."""
    _expose_lifted = False
    _lifted = None
    accepts_idx = True

    def forward(self, x, grid, idx=None):
        """``idx`` (optional int32 device tensor): the kept snapshots, when the caller has already drawn
        the bag from the NumPy stream itself (parallel.FlatTrainer does, so that the device work of a
        step can be replayed from a CUDA graph); otherwise drawn here exactly like the reference."""
        if idx is None:
            idx = _idx_tensor(draw_bag(x.shape[1], self.training), x.device)
        lifted = self.FNO_input.encode_bags(x, grid, idx=idx, pool=(self.fc0.weight.data, self.fc0.bias.data))
        if self._expose_lifted:
            # the data-parallel trainer splits backward at this tensor: once the heads' gradients are complete
            # their all-reduce runs while the per-snapshot net is still back-propagating
            self._lifted = lifted
        return self._heads(lifted)


class _NioMixin(_BagModel):
    """DeepONet(branch CNN, trunk FFN) per snapshot -> bag mean + detached fc0 -> FNO heads.

    By linearity of everything after the branch's last layer, the bag mean is taken on the
    [B, L, p] branch coefficients before the trunk contraction (exact), so the
    [B, L, n_points] DeepONet output of the reference is never materialised."""

    accepts_idx = True

    def forward(self, x, grid, idx=None):
        """``idx`` as in the NIO-FNO models: the kept snapshots as a device tensor when the caller has drawn the bag
        itself (FlatTrainer replays the whole step, cuDNN encoder included, from a CUDA graph per bag size)."""
        if idx is None:
            idx = _idx_tensor(draw_bag(x.shape[1], self.training), x.device)
        if idx is not None:
            x = x.index_select(1, idx)
        grid_flat = grid.reshape(-1, grid.shape[-1])
        u = x.unsqueeze(2) if grid.dim() == 3 else x
        coeff = self.branch(u)                                   # [B, L, p]  (cuDNN conv stack)
        basis = self.trunk(grid_flat)                            # [n_points, p]
        lifted = torch.ops.blindno_b200.deeponet_pool_contract_lift(coeff, basis, self.deeponet.b0, grid,
                                                                    self.fc0.weight.data, self.fc0.bias.data)
        return self._heads(lifted)


def make_models(variant: str):
    """Model classes of one reference directory ('1d_FPE', '1d_GPE', '2d_FPE', '2d_Non_conservative_FPE')."""
    Encoder, Encoder2D = make_encoders(variant)
    heads_1d = ("fno_V",) if variant == "1d_GPE" else ("fno_drift", "fno_diffusion")
    heads_2d = ("fno_Fx", "fno_Fy") if variant == "2d_Non_conservative_FPE" else ("fno_drift", "fno_diffusion")

    class NIOFP_FNO(_NioFnoMixin):
        head_names = heads_1d

        def __init__(self, fno_layers, width, modes, output_dim, device):
            super().__init__()
            self.device = device
            self.fno_layers = fno_layers
            self.FNO_input = FNO1d(modes=12, width=4, n_layers=2, input_dim=2, output_dim=1, device=device)
            self.fc0 = nn.Linear(2, width)
            for name in self.head_names:
                setattr(self, name, FNO1d(modes=modes, width=width, n_layers=fno_layers, input_dim=width,
                                          output_dim=1, device=device))

    class NIOFP2D_FNO(_NioFnoMixin):
        head_names = heads_2d

        def __init__(self, input_dimensions_trunk, n_hidden_layers, neurons, n_basis, fno_layers, width, modes,
                     output_dim):
            super().__init__()
            self.fno_layers = fno_layers
            self.branch = Encoder2D(n_basis)       # built, never called (kept for checkpoint compatibility)
            self.fc0 = nn.Linear(3, width)
            self.FNO_input = FNO2d(modes=12, width=4, n_layers=2, input_dim=3, output_dim=1)
            for name in self.head_names:
                setattr(self, name, FNO2d(modes=modes, width=width, n_layers=fno_layers, input_dim=width,
                                          output_dim=1))

    class _Nio1d(_NioMixin):
        def __init__(self, input_dimensions_trunk, n_hidden_layers, neurons, n_basis, fno_layers, width, modes,
                     output_dim, device):
            super().__init__()
            self.trunk = FFN(input_dimensions_trunk, n_basis, n_hidden_layers, neurons, "leaky_relu", 0.0)
            self.fno_layers = fno_layers
            self.branch = Encoder(n_basis)
            self.deeponet = DeepOnetNoBiasOrg(self.branch, self.trunk)
            self.fc0 = nn.Linear(2, width)
            self.device = device
            for name in self.head_names:
                setattr(self, name, FNO1d(modes=modes, width=width, n_layers=fno_layers, input_dim=width,
                                          output_dim=1, device=device))

    class NIOFP(_Nio1d):
        head_names = ("fno_drift", "fno_diffusion")

    class NIOFP_schrodinger(_Nio1d):
        head_names = ("fno_V",)

    class NIOFP2D(_NioMixin):
        head_names = heads_2d

        def __init__(self, input_dimensions_trunk, n_hidden_layers, neurons, n_basis, fno_layers, width, modes,
                     output_dim):
            super().__init__()
            self.trunk = FFN(input_dimensions_trunk, n_basis, n_hidden_layers, neurons, "leaky_relu", 0.0)
            self.fno_layers = fno_layers
            self.branch = Encoder2D(n_basis)
            self.deeponet = DeepOnetNoBiasOrg(self.branch, self.trunk)
            self.fc0 = nn.Linear(3, width)
            for name in self.head_names:
                setattr(self, name, FNO2d(modes=modes, width=width, n_layers=fno_layers, input_dim=width,
                                          output_dim=1))

    out = {"NIOFP_FNO": NIOFP_FNO, "NIOFP2D_FNO": NIOFP2D_FNO, "NIOFP": NIOFP, "NIOFP2D": NIOFP2D,
           "NIOFP_schrodinger": NIOFP_schrodinger, "Encoder": Encoder, "Encoder2D": Encoder2D}
    for cls in out.values():
        cls.__qualname__ = cls.__name__
    return out
