"""Lipschitz-constrained MLP of the residual flow (reference: normflows/nets/lipschitz.py:14-67 LipschitzMLP,
:132-292 InducedNormLinear with domain = codomain = 2, :642-648 Swish).

Same module tree / state_dict keys (`net.<2i>.beta`, `net.<2i+1>.{weight,bias,scale,u,v}`).  The spectral
normalisation (`compute_weight`: power iteration on the [out, in] weight, soft normalisation by max(1, sigma / coeff))
is small-matrix parameter preparation and stays in torch, exactly as in the reference; the network itself, its
Jacobian-vector and vector-Jacobian products run in libnfb200 (tensor-core GEMMs + csrc/nfb_residual.cu)."""
import math

import torch
import torch.nn.functional as F
import torch.nn.init as init
from torch import nn


class Swish(nn.Module):
    def __init__(self):
        super().__init__()
        self.beta = nn.Parameter(torch.tensor([0.5]))

    def forward(self, x):
        from .._native import swish
        return swish(x, float(F.softplus(self.beta.detach())))[0]


class InducedNormLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True, coeff=0.97, domain=2, codomain=2, n_iterations=None,
                 atol=None, rtol=None, zero_init=False, **unused_kwargs):
        super().__init__()
        if domain != 2 or codomain != 2:
            raise NotImplementedError("only the spectral norm (domain = codomain = 2) is on the CUDA path")
        self.in_features, self.out_features = in_features, out_features
        self.coeff, self.n_iterations, self.atol, self.rtol = coeff, n_iterations, atol, rtol
        self.domain, self.codomain = domain, codomain
        self.weight = nn.Parameter(torch.Tensor(out_features, in_features))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters(zero_init)
        h, w = self.weight.shape
        self.register_buffer("scale", torch.tensor(0.0))
        self.register_buffer("u", F.normalize(self.weight.new_empty(h).normal_(0, 1), p=2, dim=0))
        self.register_buffer("v", F.normalize(self.weight.new_empty(w).normal_(0, 1), p=2, dim=0))
        with torch.no_grad():
            self.compute_weight(True, n_iterations=200, atol=None, rtol=None)

    def reset_parameters(self, zero_init=False):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if zero_init:
            self.weight.data.div_(1000)
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def compute_weight(self, update=True, n_iterations=None, atol=None, rtol=None):
        """nets/lipschitz.py:221-268."""
        u, v, weight = self.u, self.v, self.weight
        if update:
            n_iterations = self.n_iterations if n_iterations is None else n_iterations
            atol = self.atol if atol is None else atol
            rtol = self.rtol if rtol is None else atol
            if n_iterations is None and (atol is None or rtol is None):
                raise ValueError("Need one of n_iteration or (atol, rtol).")
            max_itrs = 200 if n_iterations is None else n_iterations
            with torch.no_grad():
                for _ in range(max_itrs):
                    if n_iterations is None and atol is not None and rtol is not None:
                        old_v, old_u = v.clone(), u.clone()
                    u = F.normalize(torch.mv(weight, v), p=2, dim=0, out=u)
                    v = F.normalize(torch.mv(weight.t(), u), p=2, dim=0, out=v)
                    if n_iterations is None and atol is not None and rtol is not None:
                        err_u = torch.norm(u - old_u) / (u.nelement() ** 0.5)
                        err_v = torch.norm(v - old_v) / (v.nelement() ** 0.5)
                        if err_u < atol + rtol * torch.max(u) and err_v < atol + rtol * torch.max(v):
                            break
                self.v.copy_(v)
                self.u.copy_(u)
                u, v = u.clone(), v.clone()
        sigma = torch.dot(u, torch.mv(weight, v))
        with torch.no_grad():
            self.scale.copy_(sigma)
        factor = torch.max(torch.ones(1).to(weight.device), sigma / self.coeff)
        return weight / factor

    def forward(self, input):
        from .._native import linear
        with torch.no_grad():
            return linear(input, self.compute_weight(update=False), self.bias)


class LipschitzMLP(nn.Module):
    def __init__(self, channels, lipschitz_const=0.97, max_lipschitz_iter=5, lipschitz_tolerance=None, init_zeros=True):
        super().__init__()
        self.n_layers = len(channels) - 1
        self.channels, self.lipschitz_const = channels, lipschitz_const
        self.max_lipschitz_iter, self.lipschitz_tolerance, self.init_zeros = max_lipschitz_iter, lipschitz_tolerance, init_zeros
        layers = []
        for i in range(self.n_layers):
            layers += [Swish(), InducedNormLinear(in_features=channels[i], out_features=channels[i + 1],
                                                  coeff=lipschitz_const, domain=2, codomain=2,
                                                  n_iterations=max_lipschitz_iter, atol=lipschitz_tolerance,
                                                  rtol=lipschitz_tolerance,
                                                  zero_init=init_zeros if i == (self.n_layers - 1) else False)]
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x)
