"""Runs the UNMODIFIED reference (baseline/_ref/normflows, copied verbatim by __graft_entry__.build()) in a separate
process and writes a fixture .npz for shapes too large to commit as goldens.  Used by `-m gpu` tests only (the copy
travels to the GPU box; /root/reference is never read here).
    python tests/ref_runner.py glow_c3 <out.npz> [batch]
Exit code 3: baseline/_ref is absent (the calling test skips)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
if not os.path.isfile(os.path.join(REF, "normflows", "__init__.py")):
    sys.exit(3)
sys.path.insert(0, REF)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import normflows as nf  # noqa: E402

assert nf.__file__.startswith(REF)


def glow_c3(out, batch):
    """BASELINE config 3 = examples/glow.ipynb cell 2 verbatim: L=3, K=16, hidden 256, 3x32x32, 10 classes."""
    torch.manual_seed(0)
    L, K, hidden, shape, ncls = 3, 16, 256, (3, 32, 32), 10
    q0, merges, flows = [], [], []
    for i in range(L):
        flows_ = []
        for j in range(K):
            flows_ += [nf.flows.GlowBlock(shape[0] * 2 ** (L + 1 - i), hidden, split_mode="channel", scale=True)]
        flows_ += [nf.flows.Squeeze()]
        flows += [flows_]
        if i > 0:
            merges += [nf.flows.Merge()]
            ls = (shape[0] * 2 ** (L - i), shape[1] // 2 ** (L - i), shape[2] // 2 ** (L - i))
        else:
            ls = (shape[0] * 2 ** (L + 1), shape[1] // 2 ** L, shape[2] // 2 ** L)
        q0 += [nf.distributions.ClassCondDiagGaussian(ls, ncls)]
    model = nf.MultiscaleFlow(q0, flows, merges)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(batch, *shape, generator=g)
    y = torch.randint(ncls, (batch,), generator=g)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    model = model.to(dev).double()
    x, y = x.to(dev).double(), y.to(dev)
    with torch.no_grad():
        model.log_prob(x, y)  # ActNorm data-dependent init
        gp = torch.Generator().manual_seed(2)
        for p in model.parameters():  # move the zero-initialised last convolutions / base off their init
            p.add_((0.02 * torch.randn(p.shape, generator=gp, dtype=torch.float64)).to(dev))
        lp = model.log_prob(x, y)
    res = {"x": x.cpu().numpy(), "y": y.cpu().numpy(), "log_prob_f64": lp.cpu().numpy()}
    for k, v in model.state_dict().items():
        res["sd__" + k] = v.detach().cpu().numpy()
    np.savez(out, **res)


if __name__ == "__main__":
    case, out = sys.argv[1], sys.argv[2]
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    {"glow_c3": glow_c3}[case](out, batch)
