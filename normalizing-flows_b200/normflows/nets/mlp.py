"""MLP parameter container (reference: normflows/nets/mlp.py:5-58).

Same constructor and `state_dict` keys (`net.<i>.weight|bias`).  The arithmetic runs inside the
affine-stack kernel (csrc/nfb_affine.cu); calling the module directly evaluates it through the same
library as a one-layer affine problem is not needed -- it is only ever used as s/t/param_map."""
from torch import nn


class MLP(nn.Module):
    def __init__(self, layers, leaky=0.0, score_scale=None, output_fn=None, output_scale=None,
                 init_zeros=False, dropout=None):
        super().__init__()
        if output_fn is not None or score_scale is not None or output_scale is not None:
            raise NotImplementedError("MLP output_fn / scaling is not on the CUDA path")
        if dropout is not None:
            raise NotImplementedError("MLP dropout is not on the CUDA path")
        mods = []
        for k in range(len(layers) - 2):
            mods += [nn.Linear(layers[k], layers[k + 1]), nn.LeakyReLU(leaky)]
        mods.append(nn.Linear(layers[-2], layers[-1]))
        if init_zeros:
            nn.init.zeros_(mods[-1].weight)
            nn.init.zeros_(mods[-1].bias)
        self.net = nn.Sequential(*mods)
        self.leaky = leaky
        self.layer_sizes = list(layers)

    def linear_layers(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]

    def forward(self, x):
        raise RuntimeError("MLP is a parameter container on the CUDA path; it is evaluated inside the "
                           "fused affine kernels (MaskedAffineFlow / AffineCouplingBlock)")
