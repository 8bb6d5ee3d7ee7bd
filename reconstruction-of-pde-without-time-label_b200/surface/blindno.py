"""NIOModules.py surface, BlinDNO family: the permutation-invariant U-Net with attention over the bag.

  PermInvUNet_attn            2d_FPE/NIOModules.py:1086-1181, 2d_Non_conservative_FPE/NIOModules.py:932-1040
  PermInvUNet_attn1D(_bag)    1d_FPE/NIOModules.py:212-443
  PermInvUNet_attn1D_bag(_GPE)  1d_GPE/NIOModules.py:342-560

SURVEY.md section 8(f) N1: these models end in the same FNO output heads as NIO-FNO, so the heads run
through ``torch.ops.blindno_b200.fno_net`` (two heads on two streams); the per-snapshot U-Net encoder and
the bag attention are torch library calls (cuDNN / cuBLAS) -- they are outside the north-star path.
One dimension-generic implementation replaces the reference's five near-identical classes; attribute
names, creation order (so a seeded construction draws the same initial weights) and state_dict layout
follow the reference class of each directory.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from .fno import FNO1d, FNO2d
from .nio import _BagModel, _idx_tensor, draw_bag


def _nd(name: str, nd: int):
    return getattr(nn, f"{name}{nd}d")


class _ConvNeXt(nn.Module):
    """Depthwise 7-tap conv -> LayerNorm over channels -> 4x MLP with GELU -> residual."""

    def __init__(self, dim: int, nd: int):
        super().__init__()
        self.dwconv = _nd("Conv", nd)(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        y = self.dwconv(x).movedim(1, -1)                         # channels last for LayerNorm / Linear
        y = self.pwconv2(self.act(self.pwconv1(self.norm(y))))
        return x + y.movedim(-1, 1)


class _BagAttention(nn.Module):
    """Self-attention over the snapshots of a bag with the flattened feature map as the token
    (no projections), residual, LayerNorm over the features: permutation-equivariant in the bag."""

    def __init__(self, *feature_shape: int):
        super().__init__()
        self.D = math.prod(feature_shape)
        self.norm = nn.LayerNorm(self.D)

    def forward(self, x):                                         # [B, L, *feature_shape]
        tokens = x.reshape(x.shape[0], x.shape[1], self.D)
        weights = torch.softmax(tokens @ tokens.transpose(1, 2) / math.sqrt(self.D), dim=-1)
        return self.norm(weights @ tokens + tokens).reshape(x.shape)


class _PermInvUNet(_BagModel):
    """Per-snapshot U-Net encoder; at every scale the bag is mixed by _BagAttention and averaged; the decoder
    runs on the averaged maps; a 1x1 conv lifts to the FNO heads' width."""
    nd = 2
    block = "convnext"          # "convnext" | "bn_relu" (2d_Non_conservative_FPE)
    subsample = True            # draw a bag subsample in training mode (the *_bag classes and the 2-D ones)
    head_defs = ()              # (name, modes) of every FNO head that is registered, in order
    head_names = ()             # the heads whose outputs are concatenated

    def _stage(self, c_in: int, c_out: int):
        conv = _nd("Conv", self.nd)(c_in, c_out, kernel_size=3, padding=1)
        if self.block == "convnext":
            return nn.Sequential(conv, _ConvNeXt(c_out, self.nd))
        return nn.Sequential(conv, _nd("BatchNorm", self.nd)(c_out), nn.ReLU(inplace=True))

    def _build(self, in_ch, base_ch, depth, input_size, width, device=None):
        nd = self.nd
        self.depth, self.width = depth, width
        self.chs = [base_ch * 2 ** i for i in range(depth + 1)]
        size0 = tuple(input_size) if nd == 2 else (int(input_size),)
        sizes = [size0]
        for _ in range(depth):
            sizes.append(tuple(s // 2 for s in sizes[-1]))
        # a stride-2 transposed conv doubles the map; output_padding restores odd sizes on the way up
        out_pads = [tuple(big - 2 * small for big, small in zip(sizes[i], sizes[i + 1])) for i in reversed(range(depth))]

        self.down_convs, self.pools = nn.ModuleList(), nn.ModuleList()
        self.down_convs.append(self._stage(in_ch, self.chs[0]))
        for i in range(depth):
            self.pools.append(_nd("MaxPool", nd)(2))
            self.down_convs.append(self._stage(self.chs[i], self.chs[i + 1]))
        self.skip_norms = nn.ModuleList([_nd("BatchNorm", nd)(c) for c in self.chs])
        self.temp_atts = nn.ModuleList([_BagAttention(self.chs[i], *sizes[i]) for i in range(depth + 1)])
        self.up_transposes, self.up_convs = nn.ModuleList(), nn.ModuleList()
        for pad, i in zip(out_pads, reversed(range(depth))):
            self.up_transposes.append(_nd("ConvTranspose", nd)(self.chs[i + 1], self.chs[i], kernel_size=2, stride=2,
                                                                output_padding=pad if nd == 2 else pad[0]))
            self.up_convs.append(self._stage(2 * self.chs[i], self.chs[i]))
        self.final_conv = _nd("Conv", nd)(self.chs[0], width, kernel_size=1)
        for name, modes in self.head_defs:
            head = FNO2d(modes=modes, width=width, n_layers=3, input_dim=width, output_dim=1) if nd == 2 else \
                FNO1d(modes=modes, width=width, n_layers=3, input_dim=width, output_dim=1, device=device)
            setattr(self, name, head)

    fused_attention = True      # CUDA: attention + bag mean through ops.bag_attention_mean (bags of <= 128 snapshots)

    def _bag_mean(self, level: int, maps, n_bags: int):
        seq = maps.reshape(n_bags, -1, *maps.shape[1:])
        att = self.temp_atts[level]
        if self.fused_attention and seq.is_cuda and seq.shape[1] <= ops.BAG_ATTENTION_MAX_KEEP:
            out = ops.bag_attention_mean(seq.reshape(n_bags, seq.shape[1], att.D), att.norm.weight, att.norm.bias, att.norm.eps)
            return out.reshape(n_bags, *maps.shape[1:])
        return att(seq).mean(dim=1)

    accepts_idx = True

    def forward(self, x, grid=None, idx=None):
        """``model(x)`` as in the reference.  ``grid`` is accepted and ignored (these models carry no coordinate
        input) and ``idx`` is the bag drawn by the caller, as for the NIO / NIO-FNO models, so that
        ``parallel.FlatTrainer`` can replay the whole step from a CUDA graph per bag size."""
        if idx is None and self.subsample:
            idx = _idx_tensor(draw_bag(x.shape[1], self.training), x.device)
        if idx is not None:
            x = x.index_select(1, idx)
        n_bags = x.shape[0]
        h = x.reshape(n_bags * x.shape[1], 1, *x.shape[2:])
        skips = []
        for level in range(self.depth + 1):
            h = self.down_convs[level](h)
            skips.append(h)
            if level < self.depth:
                h = self.pools[level](h)
        h = self._bag_mean(self.depth, h, n_bags)
        for k, level in enumerate(reversed(range(self.depth))):
            skip = self.skip_norms[level](self._bag_mean(level, skips[level], n_bags))
            h = self.up_convs[k](torch.cat([self.up_transposes[k](h), skip], dim=1))
        return self._heads(self.final_conv(h).movedim(1, -1).contiguous())


def make_blindno_models(variant: str):
    """BlinDNO classes of one reference directory, by the names its scripts import."""
    out = {}
    if variant.startswith("2d"):
        nc = variant == "2d_Non_conservative_FPE"

        class PermInvUNet_attn(_PermInvUNet):
            nd = 2
            block = "bn_relu" if nc else "convnext"
            head_defs = (("fno_drift", 32), ("fno_diffusion", 32)) + ((("fno_Fx", 32), ("fno_Fy", 32)) if nc else ())
            head_names = ("fno_Fx", "fno_Fy") if nc else ("fno_drift", "fno_diffusion")

            def __init__(self, in_ch=1, out_ch=2, base_ch=1, depth=4, input_size=(61, 61)):
                super().__init__()
                self._build(in_ch, base_ch, depth, input_size, width=12)

        out["PermInvUNet_attn"] = PermInvUNet_attn
    elif variant == "1d_FPE":
        class PermInvUNet_attn1D(_PermInvUNet):
            nd = 1
            subsample = False
            head_defs = (("fno_drift", 15), ("fno_diffusion", 15))
            head_names = ("fno_drift", "fno_diffusion")

            def __init__(self, in_ch=1, out_ch=2, base_ch=1, depth=4, input_size=61, device=None):
                super().__init__()
                self.device = device
                self._build(in_ch, base_ch, depth, input_size, width=30, device=device)

        class PermInvUNet_attn1D_bag(PermInvUNet_attn1D):
            subsample = True

        out.update(PermInvUNet_attn1D=PermInvUNet_attn1D, PermInvUNet_attn1D_bag=PermInvUNet_attn1D_bag)
    elif variant == "1d_GPE":
        class PermInvUNet_attn1D_bag(_PermInvUNet):
            nd = 1
            head_defs = (("fno_V", 30),)
            head_names = ("fno_V",)

            def __init__(self, in_ch=1, out_ch=2, base_ch=1, depth=4, input_size=61, device=None):
                super().__init__()
                self.device = device
                self._build(in_ch, base_ch, depth, input_size, width=10, device=device)

        class PermInvUNet_attn1D_bag_GPE(_PermInvUNet):
            nd = 1
            head_names = ("fno_V",)

            def __init__(self, in_ch=1, out_ch=2, base_ch=1, depth=4, input_size=61, device=None, width=None, modes=None):
                super().__init__()
                self.device, self.modes = device, modes
                self.head_defs = (("fno_V", modes),)
                self._build(in_ch, base_ch, depth, input_size, width=width, device=device)

        out.update(PermInvUNet_attn1D_bag=PermInvUNet_attn1D_bag, PermInvUNet_attn1D_bag_GPE=PermInvUNet_attn1D_bag_GPE)
    else:
        raise ValueError(f"unknown variant {variant!r}")
    for cls in out.values():
        cls.__qualname__ = cls.__name__
    return out
