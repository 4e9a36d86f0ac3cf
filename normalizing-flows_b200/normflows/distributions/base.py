"""Base distributions: only what the density pass ends in (reference:
normflows/distributions/base.py:8-49 BaseDistribution, :53-103 DiagGaussian)."""
import numpy as np
import torch
from torch import nn

from .. import _lib as L
from .._native import require_cuda_f32


class BaseDistribution(nn.Module):
    def forward(self, num_samples=1):
        raise NotImplementedError

    def log_prob(self, z):
        raise NotImplementedError

    def sample(self, num_samples=1, **kwargs):
        z, _ = self.forward(num_samples, **kwargs)
        return z


class DiagGaussian(BaseDistribution):
    def __init__(self, shape, trainable=True):
        super().__init__()
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        self.shape, self.n_dim, self.d = shape, len(shape), int(np.prod(shape))
        if trainable:
            self.loc = nn.Parameter(torch.zeros(1, *shape))
            self.log_scale = nn.Parameter(torch.zeros(1, *shape))
        else:
            self.register_buffer("loc", torch.zeros(1, *shape))
            self.register_buffer("log_scale", torch.zeros(1, *shape))
        self.temperature = None

    def _log_scale(self):
        return self.log_scale if self.temperature is None else self.log_scale + np.log(self.temperature)

    def forward(self, num_samples=1, context=None):
        # sampling from the base is off the hot path (core.py:167-180): plain torch RNG
        eps = torch.randn((num_samples,) + self.shape, dtype=self.loc.dtype, device=self.loc.device)
        ls = self._log_scale()
        z = self.loc + torch.exp(ls) * eps
        log_p = -0.5 * self.d * np.log(2 * np.pi) - torch.sum(ls + 0.5 * eps ** 2,
                                                              list(range(1, self.n_dim + 1)))
        return z, log_p

    def log_prob(self, z, context=None):
        z = require_cuda_f32(z)
        ls = self._log_scale().contiguous()
        out = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            L.check(L.lib().nfb_diag_gaussian_log_prob(L.ptr(z), L.ptr(self.loc), L.ptr(ls), L.ptr(out),
                                                       z.shape[0], self.d, 0, L.stream_ptr()))
        return out


class ClassCondDiagGaussian(BaseDistribution):
    """Class-conditional diagonal Gaussian (reference: distributions/base.py:281-344); `log_prob(z, y)` with
    integer labels runs in csrc/nfb_glow.cu."""

    def __init__(self, shape, num_classes):
        super().__init__()
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        self.shape, self.n_dim, self.d, self.num_classes = shape, len(shape), int(np.prod(shape)), num_classes
        self.loc = nn.Parameter(torch.zeros(*shape, num_classes))
        self.log_scale = nn.Parameter(torch.zeros(*shape, num_classes))
        self.temperature = None

    def _log_scale(self):
        return self.log_scale if self.temperature is None else self.log_scale + np.log(self.temperature)

    def forward(self, num_samples=1, y=None):
        """distributions/base.py:302-325: z = loc[..., y] + exp(log_scale[..., y]) * eps and its log-density.
        The random draws (labels, eps) and the per-class parameter gather are torch device ops (plumbing; the
        reference's generator stream cannot be reproduced anyway), the density is the CUDA kernel."""
        dev = self.loc.device
        if y is not None:
            num_samples = len(y)
            if y.dim() != 1:
                y = torch.argmax(y, dim=1)
            y = y.to(device=dev, dtype=torch.int64)
        else:
            y = torch.randint(self.num_classes, (num_samples,), device=dev)
        with torch.no_grad():
            eps = torch.randn((num_samples,) + self.shape, dtype=self.loc.dtype, device=dev)
            loc = self.loc.detach().movedim(-1, 0)[y]
            log_scale = self._log_scale().detach().movedim(-1, 0)[y]
            z = (loc + torch.exp(log_scale) * eps).contiguous()
        return z, self.log_prob(z, y)

    def log_prob(self, z, y):
        z = require_cuda_f32(z)
        if y.dim() != 1:
            y = torch.argmax(y, dim=1)  # one-hot rows (base.py:336-337 accepts both)
        y = y.to(device=z.device, dtype=torch.int64).contiguous()
        ls = self._log_scale().contiguous()  # temperature annealing: log_scale + log T (base.py:318-319,339-340)
        out = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
        if z.shape[0]:
            with torch.cuda.device(z.device):
                L.check(L.lib().nfb_class_cond_diag_gaussian_log_prob(
                    L.ptr(z), L.ptr(y), L.ptr(self.loc), L.ptr(ls), L.ptr(out), z.shape[0], self.d,
                    self.num_classes, 0, L.stream_ptr()))
        return out


class ConditionalDiagGaussian(BaseDistribution):
    """Diagonal Gaussian whose mean / log-scale come from a context encoder (distributions/base.py:106-155): the
    encoder output's first half is the mean, the second half the log standard deviation.  The per-sample density is a
    row-wise kernel launch on the standardised residual."""

    def __init__(self, shape, context_encoder):
        super().__init__()
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        self.shape, self.n_dim, self.d = shape, len(shape), int(np.prod(shape))
        self.context_encoder = context_encoder

    def _params(self, context):
        enc = self.context_encoder(context)
        split = enc.shape[-1] // 2
        return enc[..., :split], enc[..., split:]

    def forward(self, num_samples=1, context=None):
        mean, log_scale = self._params(context)
        eps = torch.randn((num_samples,) + self.shape, dtype=mean.dtype, device=mean.device)
        z = mean + torch.exp(log_scale) * eps
        log_p = -0.5 * self.d * np.log(2 * np.pi) - torch.sum(log_scale + 0.5 * eps ** 2, list(range(1, self.n_dim + 1)))
        return z, log_p

    def log_prob(self, z, context=None):
        z = require_cuda_f32(z)
        mean, log_scale = self._params(context)
        # standardise per sample, then the unit-Gaussian density kernel; the log-scale term is a row sum
        u = ((z - mean) * torch.exp(-log_scale)).contiguous().reshape(z.shape[0], -1)
        zeros = torch.zeros(self.d, device=z.device)
        out = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
        if z.shape[0]:
            with torch.cuda.device(z.device):
                L.check(L.lib().nfb_diag_gaussian_log_prob(L.ptr(u), L.ptr(zeros), L.ptr(zeros), L.ptr(out), z.shape[0],
                                                           self.d, 0, L.stream_ptr()))
        return out - torch.sum(log_scale.reshape(z.shape[0], -1), dim=1)


class GlowBase(BaseDistribution):
    """Base distribution of the Glow model (reference: distributions/base.py:347-471): diagonal Gaussian with one mean
    and one log-scale per CHANNEL (`loc * exp(loc_logs * f)`, `log_scale * exp(log_scale_logs * f)`), optionally shifted
    per class (`loc_cc`, `log_scale_cc`).  The per-channel / per-class parameter tables are a few hundred numbers and are
    assembled with torch on the device (parameter preparation, like the reference); the density of the batch is the CUDA
    kernel (csrc/nfb_kernels.cu diag_gauss_kernel / csrc/nfb_glow.cu class-conditional twin)."""

    def __init__(self, shape, num_classes=None, logscale_factor=3.0):
        super().__init__()
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        self.shape, self.n_dim = shape, len(shape)
        self.num_pix = int(np.prod(shape[1:]))
        self.d = int(np.prod(shape))
        self.num_classes = num_classes
        self.class_cond = num_classes is not None
        self.logscale_factor = logscale_factor
        one = (1, shape[0]) + (1,) * (self.n_dim - 1)
        self.loc = nn.Parameter(torch.zeros(*one))
        self.loc_logs = nn.Parameter(torch.zeros(*one))
        self.log_scale = nn.Parameter(torch.zeros(*one))
        self.log_scale_logs = nn.Parameter(torch.zeros(*one))
        if self.class_cond:
            self.loc_cc = nn.Parameter(torch.zeros(num_classes, shape[0]))
            self.log_scale_cc = nn.Parameter(torch.zeros(num_classes, shape[0]))
        self.temperature = None

    def _channel_params(self):
        """([C] or [K, C]) mean and log-scale per channel (per class), base.py:397-424 / 438-461."""
        with torch.no_grad():
            loc = (self.loc * torch.exp(self.loc_logs * self.logscale_factor)).reshape(1, -1)
            ls = (self.log_scale * torch.exp(self.log_scale_logs * self.logscale_factor)).reshape(1, -1)
            if self.class_cond:
                loc = loc + self.loc_cc
                ls = ls + self.log_scale_cc
            if self.temperature is not None:
                ls = ls + np.log(self.temperature)
        return loc, ls

    @staticmethod
    def _labels(y):
        return y if y.dim() == 1 else torch.argmax(y, dim=1)   # one-hot rows select their class (base.py:403-411)

    def forward(self, num_samples=1, y=None):
        dev = self.loc.device
        loc, ls = self._channel_params()
        if self.class_cond:
            if y is not None:
                num_samples = len(y)
                y = self._labels(y).to(device=dev, dtype=torch.int64)
            else:
                y = torch.randint(self.num_classes, (num_samples,), device=dev)
            loc, ls = loc[y], ls[y]                                         # [B, C]
        view = (-1, self.shape[0]) + (1,) * (self.n_dim - 1)
        with torch.no_grad():
            eps = torch.randn((num_samples,) + self.shape, dtype=self.loc.dtype, device=dev)
            z = (loc.reshape(view) + torch.exp(ls.reshape(view)) * eps).contiguous()
        return z, self.log_prob(z, y)

    def log_prob(self, z, y=None):
        z = require_cuda_f32(z)
        loc, ls = self._channel_params()
        out = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
        if not z.shape[0]:
            return out
        with torch.cuda.device(z.device):
            if self.class_cond:
                y = self._labels(y).to(device=z.device, dtype=torch.int64).contiguous()
                # [dim, K] tables: every pixel of channel c carries the channel's value
                lt = loc.t().repeat_interleave(self.num_pix, dim=0).contiguous()
                st = ls.t().repeat_interleave(self.num_pix, dim=0).contiguous()
                L.check(L.lib().nfb_class_cond_diag_gaussian_log_prob(L.ptr(z), L.ptr(y), L.ptr(lt), L.ptr(st), L.ptr(out),
                                                                      z.shape[0], self.d, self.num_classes, 0,
                                                                      L.stream_ptr()))
            else:
                lt = loc.reshape(-1).repeat_interleave(self.num_pix).contiguous()
                st = ls.reshape(-1).repeat_interleave(self.num_pix).contiguous()
                L.check(L.lib().nfb_diag_gaussian_log_prob(L.ptr(z), L.ptr(lt), L.ptr(st), L.ptr(out), z.shape[0], self.d,
                                                           0, L.stream_ptr()))
        return out
