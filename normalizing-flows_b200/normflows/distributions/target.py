"""Synthetic 2-D target used to generate benchmark/test inputs (reference:
normflows/distributions/target.py:99-129 TwoMoons; rejection sampler :34-73)."""
import numpy as np
import torch
from torch import nn


class TwoMoons(nn.Module):
    def __init__(self):
        super().__init__()
        self.n_dims = 2
        self.max_log_prob = 0.0
        self.register_buffer("prop_scale", torch.tensor(6.0))
        self.register_buffer("prop_shift", torch.tensor(-3.0))

    def log_prob(self, z):
        a = torch.abs(z[:, 0])
        return (-0.5 * ((torch.norm(z, dim=1) - 2) / 0.2) ** 2 - 0.5 * ((a - 2) / 0.3) ** 2
                + torch.log(1 + torch.exp(-4 * a / 0.09)))

    def sample(self, num_samples=1):
        out = torch.zeros((0, 2), dtype=self.prop_scale.dtype, device=self.prop_scale.device)
        while len(out) < num_samples:
            eps = torch.rand((num_samples, 2), dtype=out.dtype, device=out.device)
            z = self.prop_scale * eps + self.prop_shift
            accept = torch.rand(num_samples, dtype=out.dtype, device=out.device) < \
                torch.exp(self.log_prob(z) - self.max_log_prob)
            out = torch.cat([out, z[accept]], 0)
        return out[:num_samples]
