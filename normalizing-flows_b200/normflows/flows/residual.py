"""Invertible residual flow (reference: normflows/flows/residual.py:12-430, after rtqichen/residual-flows).

`Residual(net)` wraps an `iResBlock`: y = x + g(x) with a Lipschitz-constrained `g` (nets.LipschitzMLP), and
log|det(I + dg/dx)| from
  * the exact 2 x 2 Jacobian for 2-D inputs in eval mode or with brute_force=True (:148-161), or
  * the (Russian-roulette) power series sum_k (-1)^(k+1)/k c_k tr(J^k) with Hutchinson's trace estimator (:163-217,
    :355-379); in training mode with neumann_grad the reference returns the Neumann-series surrogate (:368-379), and so
    does this class.
Everything numerical runs in libnfb200: g, its Jacobian-vector products (forward mode: 2 tangents for the exact path)
and its vector-Jacobian products (reverse mode: one per power-series term) are tensor-core GEMMs (csrc/nfb_gemm_tc.cu)
around the element-wise kernels of csrc/nfb_residual.cu.  Host side only draws the random truncation n (numpy, like
the reference) and the probe vector (torch.randn_like), both injectable for exact parity tests.
Gradients of the estimator w.r.t. the parameters (training of residual flows) are not on the CUDA path yet: the
methods run under no_grad."""
import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib as L
from .._native import linear, linear_t, mul_rows, require_cuda_f32, rowdot, swish
from ..nets.lipschitz import InducedNormLinear, Swish
from .base import Flow


class Residual(Flow):
    def __init__(self, net, reverse=True, reduce_memory=True, geom_p=0.5, lamb=2.0, n_power_series=None,
                 exact_trace=False, brute_force=False, n_samples=1, n_exact_terms=2, n_dist="geometric"):
        super().__init__()
        self.reverse = reverse
        self.iresblock = iResBlock(net, n_samples=n_samples, n_exact_terms=n_exact_terms, neumann_grad=reduce_memory,
                                   grad_in_forward=reduce_memory, exact_trace=exact_trace, geom_p=geom_p, lamb=lamb,
                                   n_power_series=n_power_series, brute_force=brute_force, n_dist=n_dist)

    def forward(self, z):
        if self.reverse:
            z, log_det = self.iresblock.inverse(z, 0)
        else:
            z, log_det = self.iresblock.forward(z, 0)
        return z, -log_det.view(-1)

    def inverse(self, z):
        if self.reverse:
            z, log_det = self.iresblock.forward(z, 0)
        else:
            z, log_det = self.iresblock.inverse(z, 0)
        return z, -log_det.view(-1)


class iResBlock(nn.Module):
    def __init__(self, nnet, geom_p=0.5, lamb=2.0, n_power_series=None, exact_trace=False, brute_force=False,
                 n_samples=1, n_exact_terms=2, n_dist="geometric", neumann_grad=True, grad_in_forward=False):
        super().__init__()
        self.nnet = nnet
        self.n_dist = n_dist
        self.geom_p = nn.Parameter(torch.tensor(np.log(geom_p) - np.log(1.0 - geom_p)))
        self.lamb = nn.Parameter(torch.tensor(lamb))
        self.n_samples, self.n_power_series = n_samples, n_power_series
        self.exact_trace, self.brute_force, self.n_exact_terms = exact_trace, brute_force, n_exact_terms
        self.grad_in_forward, self.neumann_grad = grad_in_forward, neumann_grad
        self.register_buffer("last_n_samples", torch.zeros(self.n_samples))
        self.register_buffer("last_firmom", torch.zeros(1))
        self.register_buffer("last_secmom", torch.zeros(1))
        # test hooks: inject the random truncation / the probe vector of the next _logdetgrad call
        self._inject_n, self._inject_eps = None, None

    # ---- the network, split into its layers ------------------------------------------------------------
    def _layers(self):
        mods = list(self.nnet.net) if hasattr(self.nnet, "net") else None
        if not mods or len(mods) % 2 or not all(isinstance(mods[i], Swish) and isinstance(mods[i + 1], InducedNormLinear)
                                                for i in range(0, len(mods), 2)):
            raise NotImplementedError("the CUDA path of Residual takes a nets.LipschitzMLP")
        return [(mods[i], mods[i + 1]) for i in range(0, len(mods), 2)]

    @torch.no_grad()
    def _run(self, x, tangents=None, keep=False):
        """g(x); optionally pushes `tangents` [nt, B, D] forward (Jacobian-vector products) and/or keeps what the
        vector-Jacobian product needs (activation derivatives, effective weights)."""
        h, t, tape = x, tangents, []
        for sw, lin in self._layers():
            a, da = swish(h, float(F.softplus(sw.beta.detach())), want_derivative=(t is not None or keep))
            w = lin.compute_weight(update=False).contiguous()
            if t is not None:
                nt, b, width = t.shape
                t = mul_rows(t, da, nt)
                t = linear(t.reshape(nt * b, width), w).reshape(nt, b, w.shape[0])
            if keep:
                tape.append((da, w))
            h = linear(a, w, lin.bias)
        return h, t, tape

    @torch.no_grad()
    def _vjp(self, v, tape):
        """v^T J for every row (reverse mode through the taped layers)."""
        for da, w in reversed(tape):
            v = linear_t(v, w)       # (v W): cotangent w.r.t. the activation
            v = mul_rows(v[None], da, 1)[0]
        return v

    # ---- reference API ---------------------------------------------------------------------------------
    def forward(self, x, logpx=None):
        if logpx is None:
            return x + self._run(require_cuda_f32(x))[0]
        g, logdetgrad = self._logdetgrad(x)
        return x + g, logpx - logdetgrad

    def inverse(self, y, logpy=None):
        x = self._inverse_fixed_point(y)
        if logpy is None:
            return x
        return x, logpy + self._logdetgrad(x)[1]

    @torch.no_grad()
    def _inverse_fixed_point(self, y, atol=1e-5, rtol=1e-5):
        y = require_cuda_f32(y)
        x, x_prev = y - self._run(y)[0], y
        i = 0
        tol = atol + y.abs() * rtol
        while not torch.all((x - x_prev) ** 2 / tol < 1):
            x, x_prev = y - self._run(x)[0], x
            i += 1
            if i > 1000:
                break
        return x

    @torch.no_grad()
    def _logdetgrad(self, x):
        x = require_cuda_f32(x)
        if (self.brute_force or not self.training) and (x.dim() == 2 and x.shape[1] == 2):
            # exact 2 x 2 Jacobian by two forward-mode tangents (residual.py:148-161)
            b = x.shape[0]
            eye = torch.zeros(2, b, 2, device=x.device)
            eye[0, :, 0] = 1.0
            eye[1, :, 1] = 1.0
            g, jt, _ = self._run(x, tangents=eye)
            out = torch.empty(b, device=x.device)
            if b:
                with torch.cuda.device(x.device):
                    L.check(L.lib().nfb_logabsdet_i_plus_j_2x2(L.ptr(jt.contiguous()), b, L.ptr(out), L.stream_ptr()))
            return g, out.view(-1, 1)
        if x.dim() != 2:
            raise NotImplementedError("the CUDA path of Residual takes [batch, features] inputs")
        if self.exact_trace:
            raise NotImplementedError("exact_trace=True is not on the CUDA path (use brute_force for 2-D inputs)")
        if self.n_dist == "geometric":
            geom_p = torch.sigmoid(self.geom_p).item()
            sample_fn = lambda m: geometric_sample(geom_p, m)
            rcdf_fn = lambda k, offset: geometric_1mcdf(geom_p, k, offset)
        elif self.n_dist == "poisson":
            lamb = self.lamb.item()
            sample_fn = lambda m: poisson_sample(lamb, m)
            rcdf_fn = lambda k, offset: poisson_1mcdf(lamb, k, offset)
        else:
            raise ValueError("unknown n_dist " + str(self.n_dist))
        draw = (lambda m: np.asarray(self._inject_n)) if self._inject_n is not None else sample_fn
        n_samples = None
        if self.training:
            if self.n_power_series is None:
                n_samples = draw(self.n_samples)
                n_power_series = max(n_samples) + self.n_exact_terms
                exact = self.n_exact_terms
                coeff_fn = lambda k: 1 / rcdf_fn(k, exact) * sum(n_samples >= k - exact) / len(n_samples)
            else:
                n_power_series = self.n_power_series
                coeff_fn = lambda k: 1.0
        else:
            n_samples = draw(self.n_samples)
            n_power_series = max(n_samples) + 20
            coeff_fn = lambda k: 1 / rcdf_fn(k, 20) * sum(n_samples >= k - 20) / len(n_samples)
        vareps = self._inject_eps if self._inject_eps is not None else torch.randn_like(x)
        vareps = require_cuda_f32(vareps)
        self._inject_n, self._inject_eps = None, None
        g, _, tape = self._run(x, keep=True)
        n_power_series = int(n_power_series)
        if self.training and self.neumann_grad:
            # neumann_logdet_estimator (:368-379): the value is the Neumann-series surrogate of the reference
            vjp, neumann = vareps, vareps.clone()
            for k in range(1, n_power_series + 1):
                vjp = self._vjp(vjp, tape)
                neumann.add_(vjp, alpha=float((-1) ** k * coeff_fn(k)))
            logdetgrad = rowdot(self._vjp(neumann, tape), vareps)
        else:
            # basic_logdet_estimator (:355-366)
            vjp, logdetgrad = vareps, None
            for k in range(1, n_power_series + 1):
                vjp = self._vjp(vjp, tape)
                logdetgrad = rowdot(vjp, vareps, c=float((-1) ** (k + 1) / k * coeff_fn(k)), out=logdetgrad)
        if self.training and self.n_power_series is None:
            self.last_n_samples.copy_(torch.tensor(n_samples).to(self.last_n_samples))
            self.last_firmom.copy_(torch.mean(logdetgrad).to(self.last_firmom))
            self.last_secmom.copy_(torch.mean(logdetgrad ** 2).to(self.last_secmom))
        return g, logdetgrad.view(-1, 1)


def geometric_sample(p, n_samples):
    return np.random.geometric(p, n_samples)


def geometric_1mcdf(p, k, offset):
    if k <= offset:
        return 1.0
    k = k - offset
    return (1 - p) ** max(k - 1, 0)


def poisson_sample(lamb, n_samples):
    return np.random.poisson(lamb, n_samples)


def poisson_1mcdf(lamb, k, offset):
    if k <= offset:
        return 1.0
    k = k - offset
    s = 1.0
    for i in range(1, k):
        s += lamb ** i / math.factorial(i)
    return 1 - np.exp(-lamb) * s
