"""The spline arithmetic the kernels use (csrc/nfb_spline.cuh), compiled for the host by
tests/native, against the reference's golden vectors -- catches formula bugs without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SO = os.path.join(ROOT, "tests", "native", "_spline_host_check.so")


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(ROOT, "tests", "native", "spline_host_check.cu")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "normalizing-flows_b200/csrc/nfb_spline.cuh"))):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", SO, src])
    return C.CDLL(SO)


def run(lib, x, uw, uh, ud, K, tail, inverse, templated):
    n = x.shape[0]
    params = np.ascontiguousarray(np.concatenate([uw, uh, ud], axis=1).astype(np.float32))
    x = np.ascontiguousarray(x.astype(np.float32))
    y, lad = np.empty(n, np.float32), np.empty(n, np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.spline_host_check(vp(x), vp(params), n, K, C.c_float(tail), C.c_float(1.0), int(inverse),
                          int(templated), vp(y), vp(lad))
    return y, lad


@pytest.mark.parametrize("templated", [0, 1])
@pytest.mark.parametrize("inverse", [0, 1])
def test_spline_formula_vs_reference_golden(hostlib, templated, inverse):
    f = np.load(os.path.join(ROOT, "tests/golden/spline_edges.npz"))
    y, lad = run(hostlib, f["x_f32"], f["uw_f32"], f["uh_f32"], f["ud_f32"], 8, 3.0, inverse, templated)
    yr, lr = f[f"y_f64_{inverse}"], f[f"lad_f64_{inverse}"]
    # knot-hit rows 8..14: see tests/test_oracle_golden.py for why the atol is loose there
    np.testing.assert_allclose(y, yr, rtol=1e-5, atol=2e-5, equal_nan=True)
    np.testing.assert_allclose(lad, lr, rtol=1e-4, atol=5e-4, equal_nan=True)
    # exact edge semantics
    x = f["x_f32"]
    assert y[2] == x[2] and lad[2] == 0 and y[3] == x[3] and lad[3] == 0   # just outside +-B: identity
    assert y[4] == 100.0 and y[5] == -1e6 and np.isnan(y[6]) and lad[6] == 0
    assert abs(y[0] - 3.0) < 1e-5 and abs(y[1] + 3.0) < 1e-5 and abs(lad[0]) < 1e-5


def test_spline_random_vs_oracle(hostlib):
    from oracle import nf_oracle as O
    rng = np.random.default_rng(0)
    for K in (4, 8, 10):
        n = 4000
        uw, uh = rng.normal(size=(n, K)) * 2, rng.normal(size=(n, K)) * 2
        ud = rng.normal(size=(n, K - 1)) * 2
        x = rng.normal(size=n) * 2.2
        for inv in (0, 1):
            y, lad = run(hostlib, x, uw, uh, ud, K, 3.0, inv, 1)
            yo, lo = O.unconstrained_rqs(x.astype(np.float32).astype(np.float64), uw.astype(np.float32).astype(np.float64),
                                         uh.astype(np.float32).astype(np.float64), ud.astype(np.float32).astype(np.float64),
                                         inverse=bool(inv), tail_bound=3.0)
            # fp32 conditioning: where the spline is very steep/flat a 1-ulp change of theta moves y by
            # ~1e-4; the reference's own fp32 path has the same spread.  Bound the bulk and the tail.
            err = np.abs(y - yo)
            assert np.mean(err < 3e-5) > 0.998 and err.max() < 1e-3
            # same story for logabsdet: compare with what the reference's own fp32 arithmetic
            # (the oracle run in float32) loses against fp64 on the same inputs
            f32 = lambda a: a.astype(np.float32)
            _, l32 = O.unconstrained_rqs(f32(x), f32(uw), f32(uh), f32(ud), inverse=bool(inv), tail_bound=3.0)
            e_ours, e_ref32 = np.abs(lad - lo), np.abs(l32 - lo)
            assert np.mean(e_ours) < 2e-5
            assert e_ours.max() < max(4 * e_ref32.max(), 1e-3)


# ---- analytic backward of the spline element (csrc/nfb_spline_bwd.cuh; SURVEY 8f-1 groundwork) ----
BWD_SO = os.path.join(ROOT, "tests", "native", "_spline_bwd_host_check.so")
LOG2E = 1.4426950408889634


@pytest.fixture(scope="module")
def bwdlib():
    src = os.path.join(ROOT, "tests", "native", "spline_bwd_host_check.cu")
    hdrs = [os.path.join(ROOT, "normalizing-flows_b200/csrc", h) for h in ("nfb_spline_bwd.cuh", "nfb_spline.cuh")]
    if not os.path.exists(BWD_SO) or os.path.getmtime(BWD_SO) < max(map(os.path.getmtime, [src] + hdrs)):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", BWD_SO, src])
    return C.CDLL(BWD_SO)


def run_bwd(lib, x, uw, uh, ud, gy, gl, tail=3.0, use_float=0):
    """uw/uh are the reference's natural-log logits; the kernel formulation takes them times log2(e)."""
    n, K = uw.shape
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x, lw, lh, ud, gy, gl = f(x), f(uw * LOG2E), f(uh * LOG2E), f(ud), f(gy), f(gl)
    y, lad, gx = np.empty(n), np.empty(n), np.empty(n)
    glw, glh, gud = np.empty((n, K)), np.empty((n, K)), np.empty((n, K - 1))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.spline_bwd_check(vp(x), vp(lw), vp(lh), vp(ud), n, C.c_double(tail), vp(gy), vp(gl), int(use_float),
                         vp(y), vp(lad), vp(gx), vp(glw), vp(glh), vp(gud))
    return y, lad, gx, glw * LOG2E, glh * LOG2E, gud  # chain rule back to the natural-log logits


def test_spline_backward_matches_reference_autograd(bwdlib):
    f = np.load(os.path.join(ROOT, "tests/golden/spline_grads.npz"))
    y, lad, gx, guw, guh, gud = run_bwd(bwdlib, f["x"], f["uw"], f["uh"], f["ud"], f["cy"], f["cl"])
    # (fp64 on both sides; the two formulations order the knot arithmetic differently, and steep bins amplify
    # the last-bit differences to ~1e-9)
    np.testing.assert_allclose(y, f["y"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(lad, f["lad"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(gx, f["gx"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(guw, f["guw"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(guh, f["guh"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(gud, f["gud"], rtol=1e-6, atol=1e-8)
    assert (np.abs(f["x"]) > 3).sum() > 50 and (np.abs(f["gud"]) > 0).sum() > 500  # tails and interior both hit
    # float instantiation (what a kernel would run): same numbers to fp32 accuracy
    y32, lad32, gx32, guw32, guh32, gud32 = run_bwd(bwdlib, f["x"], f["uw"], f["uh"], f["ud"], f["cy"], f["cl"],
                                                    use_float=1)
    scale = lambda a: np.abs(a).max()
    for got, ref in ((gx32, f["gx"]), (guw32, f["guw"]), (guh32, f["guh"]), (gud32, f["gud"])):
        assert np.mean(np.abs(got - ref) < 2e-4 * scale(ref) + 1e-5) > 0.995


def test_spline_backward_matches_finite_differences(bwdlib):
    rng = np.random.default_rng(5)
    n, K, eps = 300, 8, 1e-6
    x = rng.uniform(-2.9, 2.9, size=n)
    uw, uh, ud = rng.normal(size=(n, K)), rng.normal(size=(n, K)), rng.normal(size=(n, K - 1))
    one, zero = np.ones(n), np.zeros(n)

    def fwd(x_, uw_, uh_, ud_):
        y, lad, *_ = run_bwd(bwdlib, x_, uw_, uh_, ud_, zero, zero)
        return y, lad
    for which, gy, gl in (("y", one, zero), ("lad", zero, one)):
        _, _, gx, guw, guh, gud = run_bwd(bwdlib, x, uw, uh, ud, gy, gl)
        pick = (lambda t: t[0]) if which == "y" else (lambda t: t[1])
        fd = (pick(fwd(x + eps, uw, uh, ud)) - pick(fwd(x - eps, uw, uh, ud))) / (2 * eps)
        ok = np.abs(fd - gx) < 1e-5 * (1 + np.abs(gx))  # elements within eps of a knot change bin: skip those
        assert ok.mean() > 0.98
        for arr, g in ((uw, guw), (uh, guh), (ud, gud)):
            for k in range(arr.shape[1]):
                d = np.zeros_like(arr)
                d[:, k] = eps
                args_p = [x, uw, uh, ud]
                args_m = [x, uw, uh, ud]
                i = [id(uw), id(uh), id(ud)].index(id(arr)) + 1
                args_p[i] = arr + d
                args_m[i] = arr - d
                fd = (pick(fwd(*args_p)) - pick(fwd(*args_m))) / (2 * eps)
                assert (np.abs(fd - g[:, k]) < 1e-5 * (1 + np.abs(g[:, k]))).mean() > 0.98, (which, i, k)


@pytest.mark.parametrize("inverse", [0, 1])
def test_spline_tails_list_vs_oracle(hostlib, inverse):
    """The per-feature-tails evaluator the circular NSF layers use (csrc/nfb_spline.cuh rqs_eval_dyn with nd = K + 1,
    csrc/nfb_kernels.cu rqs_rows_tails_kernel), compiled for the host, against the oracle's restatement of
    utils/splines.py:48-57 (itself pinned to reference-minted vectors): circular and linear features, per-feature
    bounds, inputs outside a feature's interval (the reference's list branch returns 0 for those)."""
    from oracle import nf_oracle as O
    rng = np.random.default_rng(3)
    rows, feats, K = 600, 5, 8
    circ = np.array([1, 0, 1, 0, 1], dtype=np.int32)
    tail = np.array([np.pi, 2.0, np.pi, 4.0, 3.0], dtype=np.float32)
    uw, uh = rng.normal(size=(rows, feats, K)) * 1.5, rng.normal(size=(rows, feats, K)) * 1.5
    ud = rng.normal(size=(rows, feats, K + 1)) * 1.5
    x = rng.normal(size=(rows, feats)) * 2.0
    params = np.ascontiguousarray(np.concatenate([uw, uh, ud], axis=2).astype(np.float32))
    xf = np.ascontiguousarray(x.astype(np.float32))
    y, lad = np.empty((rows, feats), np.float32), np.empty((rows, feats), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    hostlib.spline_host_check_tails(vp(xf), vp(params), rows, feats, K, K + 1, vp(tail), vp(circ), C.c_float(1.0),
                                    int(inverse), vp(y), vp(lad))
    yr, lr = O.unconstrained_rqs_tails(xf.astype(np.float64), params[..., :K].astype(np.float64),
                                       params[..., K:2 * K].astype(np.float64), params[..., 2 * K:].astype(np.float64),
                                       circ.astype(bool), inverse=bool(inverse), tail_bound=tail.astype(np.float64))
    outside = np.abs(xf) > tail[None, :]
    assert outside.any() and (y[outside] == 0).all() and (lad[outside] == 0).all()
    ok = ~outside
    np.testing.assert_allclose(y[ok], yr[ok], rtol=2e-5, atol=3e-5)
    np.testing.assert_allclose(lad[ok], lr[ok], rtol=2e-4, atol=1e-3)


@pytest.mark.parametrize("inverse", [0, 1])
def test_spline_all_circular_vs_reference_golden(hostlib, inverse):
    """tails="circular" (utils/splines.py:42-47: K derivative parameters, the last knot repeats the first, identity
    outside the interval) = the nd = K mode of nfb_rqs_spline_tails: the host build of the device evaluator against vectors
    minted from the reference (tests/golden/make_golden.py spline_circular)."""
    f = np.load(os.path.join(ROOT, "tests/golden/spline_circular.npz"))
    K = 8
    n = f["x"].shape[0]
    params = np.ascontiguousarray(np.concatenate([f["uw"], f["uh"], f["ud"]], axis=1).astype(np.float32))
    xf = np.ascontiguousarray(f["x"].astype(np.float32))
    y, lad = np.empty(n, np.float32), np.empty(n, np.float32)
    tail, circ = np.array([3.0], np.float32), np.array([1], np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    # one "feature" per row: rows = n, feats = 1
    hostlib.spline_host_check_tails(vp(xf), vp(params), n, 1, K, K, vp(tail), vp(circ), C.c_float(1.0), int(inverse),
                                    vp(y), vp(lad))
    outside = np.abs(xf) > 3.0
    assert outside.any() and (y[outside] == xf[outside]).all() and (lad[outside] == 0).all()   # identity, not zero
    np.testing.assert_allclose(y, f[f"y_{inverse}"], rtol=2e-5, atol=3e-5)
    np.testing.assert_allclose(lad, f[f"lad_{inverse}"], rtol=2e-4, atol=1e-3)
