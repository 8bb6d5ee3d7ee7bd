"""Baselines.py surface: the DeepONet branch encoders of the NIO models.

These conv + train-mode BatchNorm + LeakyReLU(0.2) stacks stay on cuDNN (SURVEY.md section 8, row A9:
the BatchNorm couples every snapshot of the batch, so only the linear tail is fused into the
bag pool).  They exist here so that state_dict keys / shapes match the reference
(Baselines.py:40-52 ConvBlock, :186-249 Encoder2D, :254-287 Encoder); the per-directory
differences are the kernel sizes of the last blocks, captured in VARIANTS.
"""
from __future__ import annotations

import torch.nn as nn

# (Encoder.final_conv3 kernel, final_conv4 kernel, final_conv4 is applied, Encoder2D.convblock7_3 kernel)
VARIANTS = {
    "1d_FPE": ((1, 4), (1, 15), False, (2, 1)),
    "1d_GPE": ((1, 7), (1, 4), True, (2, 1)),
    "2d_FPE": ((1, 5), (1, 15), True, (2, 1)),
    "2d_Non_conservative_FPE": ((1, 5), (1, 15), True, (3, 2)),
}


class ConvBlock(nn.Module):
    def __init__(self, in_fea, out_fea, kernel_size=3, stride=1, padding=1, relu_slope=0.2):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Conv2d(in_fea, out_fea, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.BatchNorm2d(out_fea),
            nn.LeakyReLU(relu_slope, inplace=True),
        )

    def forward(self, x):
        return self.layers(x)


def make_encoders(variant: str):
    k3, k4, use4, k73 = VARIANTS[variant]

    class Encoder(nn.Module):
        """1-D branch: [B, L, N] -> [B, L, output_dim]."""

        def __init__(self, output_dim, dim1=64, dim2=128, dim3=256):
            super().__init__()
            self.conv1 = ConvBlock(1, dim1, kernel_size=(1, 3), stride=(1, 2), padding=(0, 1))
            self.conv2 = ConvBlock(dim1, dim2, kernel_size=(1, 3), stride=(1, 2), padding=(0, 1))
            self.conv3 = ConvBlock(dim2, dim3, kernel_size=(1, 3), stride=(1, 2), padding=(0, 1))
            self.final_conv1 = ConvBlock(dim3, dim3, kernel_size=(1, 5), stride=(1, 1), padding=(0, 1))
            self.final_conv2 = ConvBlock(dim3, dim3, kernel_size=(1, 5), stride=(1, 1), padding=(0, 0))
            self.final_conv3 = ConvBlock(dim3, dim3, kernel_size=k3, stride=(1, 1), padding=(0, 0))
            self.final_conv4 = ConvBlock(dim3, dim3, kernel_size=k4, stride=(1, 1), padding=(0, 0))
            self.linear = nn.Linear(dim3, output_dim)

        def features(self, x):
            nb, nl, n = x.shape
            y = x.reshape(nb * nl, 1, 1, n)
            stack = [self.conv1, self.conv2, self.conv3, self.final_conv1, self.final_conv2, self.final_conv3]
            if use4:
                stack.append(self.final_conv4)
            for blk in stack:
                y = blk(y)
            return y.reshape(nb, nl, -1)

        def forward(self, x):
            return self.linear(self.features(x))

    class Encoder2D(nn.Module):
        """2-D branch: [B, L, 1, nx, ny] -> [B, L, n_out]."""

        def __init__(self, n_out, dim1=64, dim2=128, dim3=256, dim4=512, dim5=512, sample_spatial=1.0, **kwargs):
            super().__init__()
            self.convblock1 = ConvBlock(1, dim1, kernel_size=(1, 7), stride=(1, 2), padding=(0, 3))
            self.convblock2_1 = ConvBlock(dim1, dim2, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
            self.convblock2_2 = ConvBlock(dim2, dim2, kernel_size=(3, 3), padding=(1, 1))
            self.convblock3_1 = ConvBlock(dim2, dim3, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
            self.convblock3_2 = ConvBlock(dim3, dim3, kernel_size=(3, 3), padding=(1, 1))
            self.convblock4_1 = ConvBlock(dim3, dim4, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
            self.convblock4_2 = ConvBlock(dim4, dim4, kernel_size=(3, 3), padding=(1, 1))
            self.convblock7_1 = ConvBlock(dim4, dim5, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
            self.convblock7_2 = ConvBlock(dim5, dim5, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
            self.convblock7_3 = ConvBlock(dim5, dim5, kernel_size=k73, padding=0)
            self.linear = nn.Linear(512, n_out)
            self.print_bool = False

        def features(self, x):
            nb, nl = x.shape[:2]
            y = x.reshape(nb * nl, *x.shape[2:])
            for blk in (self.convblock1, self.convblock2_1, self.convblock2_2, self.convblock3_1, self.convblock3_2,
                        self.convblock4_1, self.convblock4_2, self.convblock7_1, self.convblock7_2, self.convblock7_3):
                y = blk(y)
            return y.reshape(nb, nl, -1)

        def forward(self, x):
            return self.linear(self.features(x))

    return Encoder, Encoder2D
