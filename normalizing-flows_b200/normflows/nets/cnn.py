"""ConvNet2d parameter container (reference: normflows/nets/cnn.py:5-63): Conv2d / LeakyReLU stack with
`padding = k // 2`, last conv optionally zero-initialised; same `net.<i>` state_dict keys.  The convolutions
run in csrc/nfb_glow.cu (`nfb_conv2d`)."""
from torch import nn

from .. import _lib as L


class ConvNet2d(nn.Module):
    def __init__(self, channels, kernel_size, leaky=0.0, init_zeros=True, actnorm=False, weight_std=None):
        super().__init__()
        from ..utils.nn import ActNorm
        mods = []
        for i in range(len(kernel_size) - 1):
            conv = nn.Conv2d(channels[i], channels[i + 1], kernel_size[i], padding=kernel_size[i] // 2,
                             bias=(not actnorm))
            if weight_std is not None:
                conv.weight.data.normal_(mean=0.0, std=weight_std)
            mods.append(conv)
            if actnorm:  # nets/cnn.py:45-46: activation normalisation after every conv but the last
                mods.append(ActNorm((channels[i + 1],) + (1, 1)))
            mods.append(nn.LeakyReLU(leaky))
        i = len(kernel_size)
        mods.append(nn.Conv2d(channels[i - 1], channels[i], kernel_size[i - 1], padding=kernel_size[i - 1] // 2))
        if init_zeros:
            nn.init.zeros_(mods[-1].weight)
            nn.init.zeros_(mods[-1].bias)
        self.net = nn.Sequential(*mods)
        self.leaky = leaky

    def conv_layers(self):
        return [m for m in self.net if isinstance(m, nn.Conv2d)]

    def apply_native(self, x, c0, cin):
        """y = net(x[:, c0:c0+cin]) for a contiguous CUDA NCHW tensor x; returns [B, out, H, W]."""
        import torch
        B, ctot, H, W = x.shape
        from ..utils.nn import ActNorm
        mods = list(self.net)
        cur, cur_tot, cur_c0 = x, ctot, c0
        with torch.cuda.device(x.device):
            for j, conv in enumerate(mods):
                if not isinstance(conv, nn.Conv2d):
                    continue
                last = conv is mods[-1]
                an = mods[j + 1].actNorm if j + 1 < len(mods) and isinstance(mods[j + 1], ActNorm) else None
                w, b = conv.weight, conv.bias
                y = torch.empty(B, conv.out_channels, H, W, device=x.device, dtype=torch.float32)

                def run(wt, bt, act):
                    L.check(L.lib().nfb_conv2d(L.ptr(cur), cur_tot, cur_c0, L.ptr(wt), L.ptr(bt), L.ptr(y), B,
                                               conv.in_channels, H, W, conv.out_channels, conv.kernel_size[0], act,
                                               L.stream_ptr()))
                if an is not None:
                    # ActNorm after the conv = per-channel affine: folded into the conv's weights and bias.  First
                    # call: raw conv output -> data-dependent init (flows/normalization.py:19-29), then the fold.
                    if not an._done():
                        run(w, None, -1.0)
                        an._data_init(y, "forward")
                    with torch.no_grad():
                        e = torch.exp(an.s.detach().reshape(-1))
                        w = (conv.weight.detach() * e[:, None, None, None]).contiguous()
                        b = an.t.detach().reshape(-1).contiguous()
                run(w, b, -1.0 if last else float(self.leaky))
                cur, cur_tot, cur_c0 = y, conv.out_channels, 0
        return cur

    def forward(self, x):
        from .._native import require_cuda_f32
        x = require_cuda_f32(x)
        return self.apply_native(x, 0, x.shape[1])
