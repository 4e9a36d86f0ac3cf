"""Shared test helpers: build OUR modules from a golden `spec` and load the reference state_dict."""
import numpy as np
import torch

import normflows as nf


def build_layer(L):
    t = L["type"]
    if t == "AutoregressiveRationalQuadraticSpline":
        return nf.flows.AutoregressiveRationalQuadraticSpline(
            L["num_input_channels"], L["num_blocks"], L["num_hidden_channels"],
            num_bins=L.get("num_bins", 8), tail_bound=L.get("tail_bound", 3.0))
    if t == "CoupledRationalQuadraticSpline":
        return nf.flows.CoupledRationalQuadraticSpline(
            L["num_input_channels"], L["num_blocks"], L["num_hidden_channels"],
            num_bins=L.get("num_bins", 8), tail_bound=L.get("tail_bound", 3.0),
            reverse_mask=L.get("reverse_mask", False))
    if t == "LULinearPermute":
        return nf.flows.LULinearPermute(L["num_channels"])
    if t == "MaskedAffineFlow":
        return nf.flows.MaskedAffineFlow(torch.tensor([1.0, 0.0]), nf.nets.MLP([2, 4, 2]), nf.nets.MLP([2, 4, 2]))
    if t == "ActNorm":
        return nf.flows.ActNorm(L.get("shape", 2))
    if t == "AffineCouplingBlock":
        return nf.flows.AffineCouplingBlock(nf.nets.MLP(L["mlp"]), scale_map=L.get("scale_map", "exp"),
                                            split_mode=L.get("split_mode", "channel"))
    if t == "Permute":
        return nf.flows.Permute(L["num_channels"], mode=L.get("mode", "shuffle"))
    raise KeyError(t)


def build_model(spec, sd=None):
    flows = [build_layer(L) for L in spec["flows"]]
    d = spec["q0"]["shape"][0]
    trainable = sd is None or "q0.loc" in sd and False
    model = nf.NormalizingFlow(nf.distributions.DiagGaussian(d, trainable=False), flows)
    if sd is not None:
        tsd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
        # the reference registers trainable q0 params; ours may be buffers -- same keys either way
        missing, unexpected = model.load_state_dict(tsd, strict=True)
        assert not missing and not unexpected
    return model


def annotate_spec(spec, sd):
    """Fill in constructor details the golden spec leaves implicit (MLP sizes, channel counts)."""
    for i, L in enumerate(spec["flows"]):
        p = f"flows.{i}."
        if L["type"] == "AffineCouplingBlock":
            ks = sorted({int(k[len(p + "flows.1.param_map.net."):].split(".")[0]) for k in sd
                         if k.startswith(p + "flows.1.param_map.net.")})
            sizes = [sd[f"{p}flows.1.param_map.net.{ks[0]}.weight"].shape[1]]
            sizes += [sd[f"{p}flows.1.param_map.net.{k}.weight"].shape[0] for k in ks]
            L["mlp"] = sizes
        if L["type"] == "Permute":
            L["num_channels"] = spec["q0"]["shape"][0]
        if L["type"] == "ActNorm":
            L["shape"] = spec["q0"]["shape"][0]
    return spec


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-12))
