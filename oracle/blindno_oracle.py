"""CPU oracle for the BlinDNO / NIO-FNO hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a ``dict`` of parameter tensors
(keys = the reference ``state_dict`` names), what the reference's PyTorch
modules compute on the path named in BASELINE.json.  It runs on the CPU in
fp32 through ``torch.fft`` exactly like the reference does, so it is the
checker for the CUDA product path and the ``cpu_baseline`` / ``--impl
reference`` arm of ``bench.py``.  Nothing under ``blindno_b200`` (the
product) may import it.

Pinning: the reference ships no golden vectors, so this
oracle is pinned against outputs of the reference modules themselves, imported
read-only from /root/reference by ``tests/golden/make_golden.py`` and stored
as fixtures under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays them.  When /root/reference is present (build container) the tests
also compare live.

Reference lines followed (all relative to /root/reference):
  spectral_conv1d        1d_FPE/FNOModules.py:47-59   (rfft, DC*0.5, mix, irfft)
  spectral_conv2d        2d_FPE/FNOModules.py:141-178 (rfft2, two corner blocks, irfft2)
  fno1d_forward          1d_FPE/FNOModules.py:99-122
  fno2d_forward          2d_FPE/FNOModules.py:218-240
  bag_pool_lift          2d_FPE/NIOModules.py:564-575, 1d_FPE/NIOModules.py:139-149
  niofp2d_fno_forward    2d_FPE/NIOModules.py:543-581
  niofp1d_fno_forward    1d_FPE/NIOModules.py:119-155, 1d_GPE/NIOModules.py:262-289
  draw_bag               2d_FPE/NIOModules.py:548-551 (np.random.randint, then choice)
  conv_block             1d_GPE/Baselines.py:40-52    (Conv2d, train-mode BatchNorm2d, LeakyReLU(0.2))
  encoder1d_forward      1d_GPE/Baselines.py:254-287  (1d_FPE skips final_conv4: 1d_FPE/Baselines.py:279)
  encoder2d_forward      2d_FPE/Baselines.py:186-249
  ffn_forward            1d_GPE/DeepONetModules.py:155-185 (BatchNorm1d AFTER LeakyReLU(0.01))
  deeponet_forward       1d_GPE/DeepONetModules.py:142-151 ((branch @ trunk.T + b0) / sqrt(p))
  nio1d_forward          1d_GPE/NIOModules.py:192-223 (NIOFP_schrodinger), 1d_FPE/NIOModules.py:47-84 (NIOFP)
  nio2d_forward          2d_FPE/NIOModules.py:46-83 (NIOFP2D)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# bag subsampling (A7)
# ----------------------------------------------------------------------------
def draw_bag(n_snapshots: int, training: bool):
    """Indices of the snapshots kept for this step.

    Training: ``L ~ randint(50, L0)`` then ``choice(L0, L)`` with replacement,
    both from the *global* NumPy stream, in that order; eval keeps all.
    """
    if not training:
        return None
    n_keep = np.random.randint(50, n_snapshots)
    return np.random.choice(n_snapshots, n_keep)


# ----------------------------------------------------------------------------
# spectral convolutions (A1, A2)
# ----------------------------------------------------------------------------
def _spectrum_dtype(x, as_complex):
    """The reference hard-codes float32 / cfloat for the kept-mode buffer (Q2).  Only
    when the whole oracle is run in float64 as a higher-precision arbiter for the tests
    (inputs AND parameters double) does the buffer follow."""
    if x.dtype == torch.float64:
        return torch.complex128 if as_complex else torch.float64
    return torch.cfloat if as_complex else torch.float32


def _cmix_pair(xr, xi, w):
    """(b,i,k,l) complex given as two real planes, times w[i,o,k,l,2] -> two planes."""
    wr, wi = w[..., 0], w[..., 1]
    yr = torch.einsum("bikl,iokl->bokl", xr, wr) - torch.einsum("bikl,iokl->bokl", xi, wi)
    yi = torch.einsum("bikl,iokl->bokl", xr, wi) + torch.einsum("bikl,iokl->bokl", xi, wr)
    return yr, yi


def spectral_conv2d(x, w1, w2):
    """x [B,Ci,H,W] f32, w1/w2 [Ci,Co,m1,m2,2] f32 -> [B,Co,H,W] f32."""
    nb, _, nh, nw = x.shape
    n_out, m1, m2 = w1.shape[1], w1.shape[2], w1.shape[3]
    spec = torch.view_as_real(torch.fft.rfft2(x))
    # the reference builds this buffer as float32 whatever the model dtype is
    kept = torch.zeros(nb, n_out, nh, nw // 2 + 1, 2, dtype=_spectrum_dtype(x, False), device=x.device)
    lo_r, lo_i = _cmix_pair(spec[:, :, :m1, :m2, 0], spec[:, :, :m1, :m2, 1], w1)
    kept[:, :, :m1, :m2, 0], kept[:, :, :m1, :m2, 1] = lo_r, lo_i
    hi_r, hi_i = _cmix_pair(spec[:, :, nh - m1:, :m2, 0], spec[:, :, nh - m1:, :m2, 1], w2)
    kept[:, :, nh - m1:, :m2, 0], kept[:, :, nh - m1:, :m2, 1] = hi_r, hi_i
    return torch.fft.irfft2(torch.view_as_complex(kept), s=(nh, nw))


def spectral_conv2d_c64(x, w1, w2):
    """cfloat-weight variant (1d_FPE/FNOModules.py:124-161): w [Ci,Co,m1,m2] c64."""
    return spectral_conv2d(x, torch.view_as_real(w1), torch.view_as_real(w2))


def spectral_conv1d(x, w):
    """x [B,Ci,N] f32, w [Ci,Co,m] c64 -> [B,Co,N] f32.  DC bin is halved first."""
    nb, _, n = x.shape
    n_out, m = w.shape[1], w.shape[2]
    spec = torch.fft.rfft(x)
    dc_scale = torch.ones(spec.shape[-1], dtype=x.dtype, device=x.device)
    dc_scale[0] = 0.5
    spec = spec * dc_scale
    kept = torch.zeros(nb, n_out, n // 2 + 1, dtype=_spectrum_dtype(x, True), device=x.device)
    kept[:, :, :m] = torch.einsum("bil,iol->bol", spec[:, :, :m], w)
    return torch.fft.irfft(kept, n=n)


# ----------------------------------------------------------------------------
# FNO nets (A3-A6)
# ----------------------------------------------------------------------------
def pad_amount(n: int) -> int:
    """``int(round(n * 1/4))`` -- Python banker's rounding, as the reference."""
    return int(round(n * 0.25))


def _n_layers(p, prefix):
    n = 0
    while f"{prefix}conv_list.{n}.weight" in p:
        n += 1
    return n


def fno1d_forward(p, x, prefix=""):
    """x [B',N,Cin] -> [B',N,Cout].  Lift, right-pad, layers, crop, project."""
    h = F.linear(x, p[prefix + "fc0.weight"], p[prefix + "fc0.bias"]).transpose(1, 2)
    extra = pad_amount(h.shape[-1])
    h = F.pad(h, [0, extra])
    depth = _n_layers(p, prefix)
    for k in range(depth):
        spec = spectral_conv1d(h, p[f"{prefix}spectral_list.{k}.weights1"])
        pw = F.conv1d(h, p[f"{prefix}conv_list.{k}.weight"], p[f"{prefix}conv_list.{k}.bias"])
        h = spec + pw
        if k + 1 < depth:
            h = F.gelu(h)
    h = h[..., : h.shape[-1] - extra].transpose(1, 2)
    h = F.gelu(F.linear(h, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def fno2d_forward(p, x, prefix=""):
    """x [B',H,W,Cin] -> [B',H',W',1] (H'=H, W'=W on square grids)."""
    h = F.linear(x, p[prefix + "fc0.weight"], p[prefix + "fc0.bias"]).permute(0, 3, 1, 2)
    extra_w = pad_amount(h.shape[-1])
    extra_h = pad_amount(h.shape[-2])
    h = F.pad(h, [0, extra_w, 0, extra_h])
    depth = _n_layers(p, prefix)
    for k in range(depth):
        w1 = p[f"{prefix}spectral_list.{k}.weights1"]
        w2 = p[f"{prefix}spectral_list.{k}.weights2"]
        spec = spectral_conv2d_c64(h, w1, w2) if w1.is_complex() else spectral_conv2d(h, w1, w2)
        pw = F.conv2d(h, p[f"{prefix}conv_list.{k}.weight"], p[f"{prefix}conv_list.{k}.bias"])
        h = spec + pw
        if k + 1 < depth:
            h = F.gelu(h)
    # the reference crops H by the W-derived amount and W by the H-derived one
    h = h[..., : h.shape[-2] - extra_w, : h.shape[-1] - extra_h].permute(0, 2, 3, 1)
    h = F.gelu(F.linear(h, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


# ----------------------------------------------------------------------------
# bag mean + lift (A8)
# ----------------------------------------------------------------------------
def bag_pool_lift(s, grid_cl, w0, b0):
    """s [B,L,*g] per-snapshot scalars, grid_cl [*g,d] -> [B,*g,width].

    ``fc0([grid, mean_L s])`` written as one matmul against
    ``[W_grid | (w_s / L) repeated L times]``; w0/b0 are detached (``.data``)
    so no gradient reaches them.
    """
    w0, b0 = w0.detach(), b0.detach()
    nb, n_keep = s.shape[0], s.shape[1]
    d = grid_cl.shape[-1]
    gdims = tuple(range(grid_cl.dim() - 1))
    g = grid_cl.permute(grid_cl.dim() - 1, *gdims).unsqueeze(0).expand(nb, *([-1] * grid_cl.dim()))
    stacked = torch.cat((g, s), dim=1)                       # [B, d+L, *g]
    wide = torch.cat([w0[:, :d], w0[:, d].reshape(-1, 1).repeat(1, n_keep) / n_keep], dim=1)
    stacked = stacked.permute(0, *range(2, stacked.dim()), 1)  # [B, *g, d+L]
    return torch.matmul(stacked, wide.T) + b0


# ----------------------------------------------------------------------------
# NIO-FNO models (A11)
# ----------------------------------------------------------------------------
def niofp2d_fno_forward(p, x, grid, heads=("fno_drift", "fno_diffusion"), idx=None):
    """x [B,L0,H,W], grid [H,W,2] -> [B,H,W,len(heads)].  idx = draw_bag(...)."""
    if idx is not None:
        x = x[:, idx]
    nb, n_keep, nh, nw = x.shape
    snap = x.reshape(nb * n_keep, nh, nw, 1)
    g = grid.unsqueeze(0).expand(nb * n_keep, nh, nw, 2)
    s = fno2d_forward(p, torch.cat((snap, g), dim=-1), "FNO_input.")
    s = s.reshape(nb, n_keep, nh, nw)
    lifted = bag_pool_lift(s, grid, p["fc0.weight"], p["fc0.bias"])
    return torch.cat([fno2d_forward(p, lifted, h + ".") for h in heads], dim=-1)


def niofp1d_fno_forward(p, x, grid, heads=("fno_drift", "fno_diffusion"), idx=None):
    """x [B,L0,N], grid [N,1] -> [B,N,len(heads)]."""
    if idx is not None:
        x = x[:, idx]
    nb, n_keep, n = x.shape
    snap = x.reshape(nb * n_keep, n, 1)
    g = grid.unsqueeze(0).expand(nb * n_keep, n, 1)
    s = fno1d_forward(p, torch.cat((snap, g), dim=-1), "FNO_input.")
    s = s.reshape(nb, n_keep, n)
    lifted = bag_pool_lift(s, grid, p["fc0.weight"], p["fc0.bias"])
    return torch.cat([fno1d_forward(p, lifted, h + ".") for h in heads], dim=-1)


# ----------------------------------------------------------------------------
# a whole train step on the CPU (bench.py cpu_baseline / --impl reference)
# ----------------------------------------------------------------------------
# ----------------------------------------------------------------------------
# NIO: DeepONet(branch CNN, trunk FFN) per snapshot -> bag mean + detached fc0 -> FNO heads  (A9, A10)
# ----------------------------------------------------------------------------
def conv_block(p, prefix, x, stride, padding, training, slope=0.2):
    """ConvBlock: Conv2d -> BatchNorm2d (batch statistics in training mode, running ones in eval) ->
    LeakyReLU(0.2).  The running statistics are not updated here (they are not an output of the path)."""
    y = F.conv2d(x, p[prefix + "layers.0.weight"], p[prefix + "layers.0.bias"], stride=stride, padding=padding)
    y = F.batch_norm(y, p[prefix + "layers.1.running_mean"].clone(), p[prefix + "layers.1.running_var"].clone(),
                     p[prefix + "layers.1.weight"], p[prefix + "layers.1.bias"], training=training,
                     momentum=0.1, eps=1e-5)
    return F.leaky_relu(y, slope)


def encoder1d_forward(p, x, prefix, training, use_final_conv4=True):
    """Encoder: [B, L, N] -> [B, L, n_basis]."""
    nb, nl, n = x.shape
    y = x.reshape(nb * nl, 1, 1, n)
    for name in ("conv1", "conv2", "conv3"):
        y = conv_block(p, f"{prefix}{name}.", y, (1, 2), (0, 1), training)
    y = conv_block(p, f"{prefix}final_conv1.", y, (1, 1), (0, 1), training)
    y = conv_block(p, f"{prefix}final_conv2.", y, (1, 1), (0, 0), training)
    y = conv_block(p, f"{prefix}final_conv3.", y, (1, 1), (0, 0), training)
    if use_final_conv4:
        y = conv_block(p, f"{prefix}final_conv4.", y, (1, 1), (0, 0), training)
    y = y.reshape(nb, nl, -1)
    return F.linear(y, p[prefix + "linear.weight"], p[prefix + "linear.bias"])


def encoder2d_forward(p, x, prefix, training):
    """Encoder2D: [B, L, 1, nx, ny] -> [B, L, n_basis]."""
    nb, nl = x.shape[:2]
    y = x.reshape(nb * nl, *x.shape[2:])
    y = conv_block(p, prefix + "convblock1.", y, (1, 2), (0, 3), training)
    for name, stride in (("2_1", 2), ("2_2", 1), ("3_1", 2), ("3_2", 1), ("4_1", 2), ("4_2", 1), ("7_1", 2), ("7_2", 2)):
        y = conv_block(p, f"{prefix}convblock{name}.", y, (stride, stride), (1, 1), training)
    y = conv_block(p, prefix + "convblock7_3.", y, (1, 1), (0, 0), training)
    y = y.reshape(nb, nl, -1)
    return F.linear(y, p[prefix + "linear.weight"], p[prefix + "linear.bias"])


def ffn_forward(p, x, prefix, training):
    """FFN trunk: LeakyReLU(0.01)(input) -> per hidden layer BatchNorm1d(LeakyReLU(Linear)) -> Linear."""
    y = F.leaky_relu(F.linear(x, p[prefix + "input_layer.weight"], p[prefix + "input_layer.bias"]), 0.01)
    k = 0
    while f"{prefix}hidden_layers.{k}.weight" in p:
        y = F.leaky_relu(F.linear(y, p[f"{prefix}hidden_layers.{k}.weight"], p[f"{prefix}hidden_layers.{k}.bias"]), 0.01)
        y = F.batch_norm(y, p[f"{prefix}batch_layers.{k}.running_mean"].clone(),
                         p[f"{prefix}batch_layers.{k}.running_var"].clone(), p[f"{prefix}batch_layers.{k}.weight"],
                         p[f"{prefix}batch_layers.{k}.bias"], training=training, momentum=0.1, eps=1e-5)
        k += 1
    return F.linear(y, p[prefix + "output_layer.weight"], p[prefix + "output_layer.bias"])


def deeponet_forward(p, coeff, basis):
    """(weights @ basis.T + b0) / sqrt(p), p = number of basis functions."""
    return (torch.matmul(coeff, basis.T) + p["deeponet.b0"]) / basis.shape[1] ** 0.5


def nio1d_forward(p, x, grid, heads=("fno_V",), idx=None, training=False, use_final_conv4=True):
    """NIOFP_schrodinger / NIOFP: x [B, L0, N], grid [N, 1] -> [B, N, len(heads)]."""
    if idx is not None:
        x = x[:, torch.as_tensor(idx)]
    coeff = encoder1d_forward(p, x, "branch.", training, use_final_conv4)
    basis = ffn_forward(p, grid, "trunk.", training)
    s = deeponet_forward(p, coeff, basis)                        # [B, L, N]
    lifted = bag_pool_lift(s, grid, p["fc0.weight"], p["fc0.bias"])
    outs = [fno1d_forward(p, lifted, prefix=h + ".") for h in heads]
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=-1)


def nio2d_forward(p, x, grid, heads=("fno_drift", "fno_diffusion"), idx=None, training=False):
    """NIOFP2D: x [B, L0, nx, ny], grid [nx, ny, 2] -> [B, nx, ny, len(heads)]."""
    if idx is not None:
        x = x[:, torch.as_tensor(idx)]
    nx, ny = grid.shape[:2]
    coeff = encoder2d_forward(p, x.unsqueeze(2), "branch.", training)
    basis = ffn_forward(p, grid.reshape(-1, 2), "trunk.", training)
    s = deeponet_forward(p, coeff, basis).reshape(x.shape[0], x.shape[1], nx, ny)
    lifted = bag_pool_lift(s, grid, p["fc0.weight"], p["fc0.bias"])
    outs = [fno2d_forward(p, lifted, prefix=h + ".") for h in heads]
    return torch.cat(outs, dim=-1)


def trainable(p):
    """The tensors Adam actually updates: everything reached by autograd.
    ``fc0.*`` is detached and the unused ``branch.*`` never gets a grad."""
    return [v for k, v in p.items() if not (k.startswith("fc0.") or k.startswith("branch."))]


def train_step(p, opt, forward, x, grid, target, **kw):
    opt.zero_grad(set_to_none=True)
    idx = draw_bag(x.shape[1], True)
    pred = forward(p, x, grid, idx=idx, **kw)
    loss = F.mse_loss(pred, target)
    loss.backward()
    opt.step()
    return loss
