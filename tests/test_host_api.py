"""CPU-side checks of the drop-in boundary: module trees / state_dict keys identical to the
reference's (taken from the golden fixtures, which hold the reference `state_dict()`), MADE masks
bit-identical, C-ABI library loads and exports every symbol the header declares, and the product
path refuses to run without CUDA (no silent fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from helpers import annotate_spec, build_model

import normflows as nf
from normflows import _lib

CASES = ["nsf_ar_d64_h256_l2", "nsf_ar_d5_h128_l3", "nsf_ar_d2_h32_l2_k4", "nsf_coupled_d64_h256_l2",
         "nsf_coupled_d5_h128_l3", "nsf_coupled_d2_h32_l2_k4", "realnvp2d", "affine_block2d", "affine_block6d"]


@pytest.mark.parametrize("name", CASES)
def test_state_dict_keys_and_shapes_match_reference(name):
    spec, sd, _ = load_golden(name)
    model = build_model(annotate_spec(spec, sd))
    ours = model.state_dict()
    assert set(ours.keys()) == set(sd.keys())
    for k, v in sd.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
    # loading the reference checkpoint verbatim works (strict)
    build_model(spec, sd)


def test_made_masks_and_degrees_bit_identical():
    for name in ("nsf_ar_d64_h256_l2", "nsf_ar_d5_h128_l3", "nsf_ar_d2_h32_l2_k4"):
        spec, sd, _ = load_golden(name)
        model = build_model(spec)  # fresh construction, masks computed by OUR MaskedLinear
        for k, v in model.state_dict().items():
            if k.endswith(".mask") or k.endswith(".degrees"):
                np.testing.assert_array_equal(v.numpy(), sd[k], err_msg=k)


def test_coupled_feature_split_matches_reference():
    spec, sd, _ = load_golden("nsf_coupled_d5_h128_l3")
    model = build_model(spec)
    for k, v in model.state_dict().items():
        if k.endswith("identity_features") or k.endswith("transform_features"):
            np.testing.assert_array_equal(v.numpy(), sd[k], err_msg=k)


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "nfb200.h")).read()
    declared = set(re.findall(r"\b(nfb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nfb_abi_version() == 1


def test_no_cpu_fallback():
    spec, sd, a = load_golden("nsf_ar_d2_h32_l2_k4")
    model = build_model(spec, sd)
    x = torch.from_numpy(a["x"].astype(np.float32))
    with pytest.raises(RuntimeError, match="no CPU/eager fallback"):
        model.log_prob(x)
    with pytest.raises(RuntimeError, match="no CPU/eager fallback"):
        model.flows[0].inverse(x)
    if not torch.cuda.is_available():
        import ctypes as C
        h = C.c_void_p()
        assert _lib.lib().nfb_flow_create(C.byref(h), 4) != 0  # no device -> loud failure
        assert b"CUDA" in _lib.lib().nfb_last_error() or b"cuda" in _lib.lib().nfb_last_error()


def test_constructor_errors_mirror_reference():
    with pytest.raises(ValueError, match="Minimal bin width too large"):
        nf.flows.AutoregressiveRationalQuadraticSpline(4, 1, 16, num_bins=2000)
    with pytest.raises(NotImplementedError):
        nf.flows.AffineCouplingBlock(nf.nets.MLP([1, 4, 2]), scale_map="tanh")
    with pytest.raises(NotImplementedError):
        nf.flows.Permute(4, mode="rotate")


def test_permute_index_lists():
    p = nf.flows.Permute(5, mode="swap")
    f, i = p._index_lists()
    z = np.arange(5)
    assert list(z[f]) == [2, 3, 4, 0, 1]      # cat(z[2:], z[:2])   (mixing.py:34-37)
    assert list(z[f][i]) == [0, 1, 2, 3, 4]   # inverse undoes it   (mixing.py:47-50)


def test_shard_rows_cover():
    from normflows.parallel import shard_rows
    for n in (0, 1, 7, 64, 65536 + 3):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_glow_multiscale_state_dict_matches_reference():
    from helpers_glow import build_glow_small
    spec, sd, _ = load_golden("glow_small")
    m = build_glow_small()
    ours = m.state_dict()
    assert set(ours) == set(sd)
    for k, v in sd.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
    build_glow_small(sd)  # strict load of the reference checkpoint


def test_flow_handle_tensor_slots_follow_module_state():
    """FlowHandle remembers WHERE each parameter lives ((dict, key) slots) instead of walking the module tree on
    every call; the refresh must still see .to()/dtype casts, load_state_dict, re-registered parameters, and --
    within 256 calls -- a swapped-out sub-module object."""
    import torch
    import normflows as nf
    torch.manual_seed(0)
    fl = []
    for i in range(3):
        fl += [nf.flows.AutoregressiveRationalQuadraticSpline(6, 1, 16), nf.flows.LULinearPermute(6)]
    fl += [nf.flows.CoupledRationalQuadraticSpline(6, 1, 16)]
    m = nf.NormalizingFlow(nf.distributions.DiagGaussian(6, trainable=False), fl)
    h = m._stack()
    slow = h._build_slots()
    assert h._slots is not None and len(slow) == len(h._slots) > 40
    same = lambda a, b: len(a) == len(b) and all(x is y for x, y in zip(a, b))
    assert same(h._tensors(), slow)
    # dtype round trip replaces every buffer object and the parameters' storage
    m.double()
    m.float()
    assert same(h._tensors(), h._tensors_slow())
    # a re-registered parameter is picked up by the dict lookup
    lin = m.flows[1].linear
    lin.bias = torch.nn.Parameter(torch.ones_like(lin.bias))
    assert same(h._tensors(), h._tensors_slow()) and any(t is lin.bias for t in h._tensors())
    # in-place updates (optimizer step, load_state_dict) keep the objects and bump _version
    v0 = lin.bias._version
    with torch.no_grad():
        lin.bias.add_(1.0)
    assert lin.bias._version == v0 + 1
    # a swapped sub-module OBJECT is invisible to the slots until the periodic slow-path check (every 256th call)
    old_net = m.flows[6].prqct.transform_net
    m.flows[6].prqct.transform_net = type(old_net)(old_net.initial_layer.in_features, old_net.final_layer.out_features,
                                                  hidden_features=16, num_blocks=1)
    for _ in range(256):
        ts = h._tensors()
    assert same(ts, h._tensors_slow())


@pytest.mark.parametrize("tag", ["cc_s", "cc_t", "ca_s", "ca_t", "gb_plain", "gb_cc"])
def test_round2b_modules_match_reference_state_dicts(tag):
    """Circular NSF layers and GlowBase: parameter / buffer names and shapes equal the reference's (the goldens hold the
    reference `state_dict()`), reference checkpoints load strict=True, and -- without CUDA -- the layers refuse to run
    (no CPU fallback)."""
    if tag.startswith("gb"):
        f = np.load(os.path.join(ROOT, "tests/golden/glow_base.npz"))
        pre = tag[3:] + "__"
        m = nf.distributions.GlowBase((4, 3, 3), num_classes=5 if tag == "gb_cc" else None)
    else:
        f = np.load(os.path.join(ROOT, "tests/golden/circular.npz"))
        pre = tag + "__"
        tbt = torch.from_numpy(np.asarray(f["tail_bound_tensor"]))
        m = {"cc_s": lambda: nf.flows.CircularCoupledRationalQuadraticSpline(6, 2, 32, [0, 2, 5], tail_bound=3.0),
             "cc_t": lambda: nf.flows.CircularCoupledRationalQuadraticSpline(6, 1, 32, [0, 2, 5], tail_bound=tbt.clone(),
                                                                              reverse_mask=True),
             "ca_s": lambda: nf.flows.CircularAutoregressiveRationalQuadraticSpline(6, 2, 32, [1, 3], tail_bound=3.0),
             "ca_t": lambda: nf.flows.CircularAutoregressiveRationalQuadraticSpline(6, 1, 32, [0, 2, 5], tail_bound=tbt.clone(),
                                                                                    permute_mask=False)}[tag]()
    ref = {k[len(pre):]: np.asarray(f[k]) for k in f.files if k.startswith(pre)}
    ours = m.state_dict()
    assert set(ours.keys()) == set(ref.keys())
    for k, v in ref.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ref.items()}, strict=True)
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            if tag.startswith("gb"):
                m.log_prob(torch.zeros(2, 4, 3, 3), torch.zeros(2, dtype=torch.long)) if tag == "gb_cc" \
                    else m.log_prob(torch.zeros(2, 4, 3, 3))
            else:
                m.inverse(torch.zeros(3, 6))
