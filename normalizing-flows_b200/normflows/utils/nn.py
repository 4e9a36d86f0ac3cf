"""normflows/utils/nn.py: small nn.Module helpers.  `ActNorm` (:25-43) wraps flows.ActNorm for use INSIDE a
conditioner (ConvNet2d(actnorm=True)); on the CUDA path ConvNet2d folds it into the preceding convolution."""
import torch
from torch import nn


class ConstScaleLayer(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.register_buffer("scale", torch.tensor(scale))

    def forward(self, input):
        return input * self.scale


class ActNorm(nn.Module):
    def __init__(self, shape):
        super().__init__()
        from ..flows.affine import ActNorm as _FlowActNorm
        self.actNorm = _FlowActNorm(shape)

    def forward(self, input):
        """y = input * exp(s) + t (broadcast over the batch dims), data-dependent init on the first call
        (flows/normalization.py:19-29).  Stand-alone use only; ConvNet2d folds the layer into its convolution."""
        an = self.actNorm
        if not an._done():
            an._data_init(input, "forward")
        return input * torch.exp(an.s) + an.t


class ClampExp(nn.Module):
    def forward(self, x):
        return torch.clamp(torch.exp(x), max=1.0)
