"""CPU suite: the arithmetic the GPU kernels execute per row, compiled for the host.

stan_b200/csrc/glm_link.cuh holds link<> / link_ext<> (log-density term, residual, d/dphi term of every family) and
their special functions; the kernels include it as device code and tests/host/link_host.cpp compiles THE SAME
SOURCE with g++.  Each row's (lp_i, r_i, x_i) is checked against the CPU oracle evaluated on a one-row, one-column
model (X = [[eta]], alpha = 0, beta = 1): value and every gradient entry of that model are closed forms of the
row's terms.  This is a regression test of the kernel math that needs no GPU; the parity tests proper (whole
kernels, through the C ABI) are the -m gpu suite.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err
from oracle.oracle import PortOracle

SD = 2.5                       # prior_alpha_sd = prior_beta_sd (oracle defaults)
LOC, SCALE = 1.0, 2.0          # prior on sigma | phi (oracle defaults)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    so = tmp_path_factory.mktemp("link_host") / "liblink_host.so"
    subprocess.run([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(so),
                    os.path.join(ROOT, "tests", "host", "link_host.cpp")], check=True, capture_output=True)
    L = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    L.link_rows.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_int, C.c_int, dp, dp, dp]
    L.digamma_host.argtypes = [C.c_double]
    L.digamma_host.restype = C.c_double
    ip = C.POINTER(C.c_int)
    L.ordered_logistic_rows.argtypes = [C.c_int, dp, ip, C.c_int, dp, dp]
    L.ordered_logistic_rows.restype = None
    L.categorical_logit_rows.argtypes = [C.c_int, C.c_int, ip, dp, dp]
    L.categorical_logit_rows.restype = None
    L.link_pair_rows.argtypes = [C.c_int, C.c_int, dp, dp, C.c_double, dp, dp, dp, dp]
    L.fm_exp_rows.argtypes = [C.c_int, dp, dp]
    L.fm_exp_rows.restype = None
    L.fm_log1p_rows.argtypes = [C.c_int, dp, dp, dp]
    L.fm_log1p_rows.restype = None
    return L


def link_rows(L, fam, eta, y, aux=None, scale=1.0):
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    n = eta.size
    lp, r, x = np.empty(n), np.empty(n), np.empty(n)
    assert L.link_rows(fam, n, dp(eta), dp(y), dp(aux) if aux is not None else None, scale, 1, 1, dp(lp), dp(r), dp(x)) == 0
    return lp, r, x


def one_row_oracle(family, eta, y, scale=None, **kw):
    """lp and gradient of the one-row model at alpha = 0, beta = 1 [, log scale]."""
    po = PortOracle(family, np.array([[eta]]), np.array([y]), **kw)
    th = [0.0, 1.0] + ([np.log(scale)] if scale is not None else [])
    return po.log_prob_grad(np.array(th), propto=True, jacobian=False)


PRIOR_B = -0.5 * (1.0 / SD) ** 2      # beta = 1 prior term under propto; alpha = 0 contributes 0


def test_digamma(host_lib):
    from scipy.special import digamma
    xs = np.concatenate([np.logspace(-6, 3, 400), np.arange(1, 40) + 0.0, [1e-300, 0.5, 11.999, 12.0, 12.001, 1e8]])
    got = np.array([host_lib.digamma_host(float(v)) for v in xs])
    ref = digamma(xs)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) < 5e-15


@pytest.mark.parametrize("family,fam", [("bernoulli_logit", 0), ("poisson_log", 1), ("binomial_logit", 3)])
def test_rows_without_scale(host_lib, family, fam):
    rng = np.random.default_rng(fam)
    eta = np.concatenate([rng.normal(0, 2, 150), [0.0, 19.99, 20.01, -19.99, -20.01, 35.0, -35.0, 1e-9, -1e-9]])
    if fam == 0:
        y, aux, kw = rng.integers(0, 2, eta.size).astype(float), None, [{} for _ in eta]
    elif fam == 1:
        eta = np.clip(eta, -30, 6)
        y, aux, kw = rng.poisson(3.0, eta.size).astype(float), None, [{} for _ in eta]
    else:
        aux = rng.integers(0, 50, eta.size).astype(float)
        y = np.floor(rng.random(eta.size) * (aux + 1))
        kw = [{"trials": np.array([int(t)])} for t in aux]
    lp, r, _ = link_rows(host_lib, fam, eta, y, aux)
    for i in range(eta.size):
        lp_o, g_o = one_row_oracle(family, eta[i], int(y[i]), **kw[i])
        assert rel_err(lp[i] + PRIOR_B, lp_o) < 1e-13, (family, eta[i], y[i], lp[i] + PRIOR_B, lp_o)
        scale = max(abs(g_o[0]), abs(g_o[1]), 1e-300)
        assert abs(r[i] - g_o[0]) <= 1e-13 * scale, (family, eta[i], y[i])                    # d/dalpha = r_i
        assert abs(eta[i] * r[i] - 1.0 / SD ** 2 - g_o[1]) <= 1e-13 * max(scale, 1.0)         # d/dbeta = eta r_i + prior


def test_rows_normal_id(host_lib):
    rng = np.random.default_rng(2)
    eta, y = rng.normal(0, 2, 120), rng.normal(0, 3, 120)
    for sigma in (0.3, 1.0, 7.5):
        z2, r, _ = link_rows(host_lib, 2, eta, y, None, sigma)      # lp_i carries z^2 (normal_id_glm_lpdf.hpp:213)
        for i in range(eta.size):
            lp_o, g_o = one_row_oracle("normal_id", eta[i], y[i], sigma)
            prior_s = -0.5 * ((sigma - LOC) / SCALE) ** 2
            assert rel_err(-0.5 * z2[i] - np.log(sigma) + PRIOR_B + prior_s, lp_o) < 1e-13
            assert abs(r[i] - g_o[0]) <= 1e-13 * max(abs(g_o[0]), 1e-300)
            dsig = (z2[i] - 1.0) / sigma - (sigma - LOC) / SCALE ** 2                          # :181-183 + prior
            assert abs(dsig * sigma - g_o[2]) <= 1e-12 * max(abs(g_o[2]), 1.0)


def test_rows_neg_binomial_2_log(host_lib):
    from scipy.special import gammaln
    rng = np.random.default_rng(4)
    eta = np.concatenate([np.clip(rng.normal(0.5, 1.5, 160), -25, 8), [0.0, -40.0, 9.0]])
    y = np.concatenate([rng.poisson(4.0, 150), [0, 1, 15, 16, 17, 18, 40, 500, 100000, 2, 3, 0, 7]]).astype(float)
    for phi in (1e-3, 0.4, 2.0, 35.0, 1e4):
        lp, r, x = link_rows(host_lib, 4, eta, y, None, phi)
        for i in range(eta.size):
            lp_o, g_o = one_row_oracle("neg_binomial_2_log", eta[i], int(y[i]), phi)
            prior_p = -0.5 * ((phi - LOC) / SCALE) ** 2
            const = phi * np.log(phi) - gammaln(phi)                # the N (phi log phi - lgamma phi) term, N = 1
            ref_lp = lp_o - PRIOR_B - prior_p - const
            # (the one-row model's lp is dominated by the phi prior for large phi: scale the bound by what was subtracted)
            assert abs(lp[i] - ref_lp) <= 2e-13 * max(abs(ref_lp), abs(const), abs(prior_p), 1.0), (phi, eta[i], y[i], lp[i], ref_lp)
            sc = max(abs(g_o[0]), abs(g_o[1]), 1.0)
            assert abs(r[i] - g_o[0]) <= 1e-12 * sc, (phi, eta[i], y[i])
            dphi = (x[i] - (phi - LOC) / SCALE ** 2) * phi          # chain through phi = exp(u), no Jacobian
            assert abs(dphi - g_o[2]) <= 1e-11 * max(abs(g_o[2]), phi * 1e-2, 1.0), (phi, eta[i], y[i], dphi, g_o[2])


def test_rows_ordered_logistic(host_lib):
    """ordered_logistic_row (not used by a kernel yet) against the oracle's ordered_logistic model on one-row data:
    theta = [beta = 1, unconstrained cut-points]; lp, d/dbeta = loc * w and the cut-point partials d1 / d2."""
    rng = np.random.default_rng(6)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for Cn in (2, 3, 6):
        u = np.concatenate([[-1.2], np.log(rng.uniform(0.3, 1.5, Cn - 2))]) if Cn > 2 else np.array([-0.4])
        cuts = np.concatenate([[u[0]], u[0] + np.cumsum(np.exp(u[1:]))])
        loc = np.concatenate([rng.normal(0, 2.5, 90), [0.0, 30.0, -30.0, cuts[0], cuts[-1]]])
        cls = rng.integers(1, Cn + 1, loc.size).astype(np.int32)
        out = np.empty(4 * loc.size)
        host_lib.ordered_logistic_rows(loc.size, dp(loc), cls.ctypes.data_as(C.POINTER(C.c_int)), Cn, dp(cuts), dp(out))
        out = out.reshape(-1, 4)
        for i in range(loc.size):
            po = PortOracle("ordered_logistic", np.array([[loc[i]]]), np.array([cls[i]]), n_classes=Cn)
            lp_o, g_o = po.log_prob_grad(np.concatenate([[1.0], u]), propto=True, jacobian=False)
            prior = PRIOR_B - 0.5 * np.sum((cuts / SD) ** 2)
            assert abs(out[i, 0] + prior - lp_o) <= 1e-13 * max(abs(lp_o), 1.0), (Cn, loc[i], cls[i])
            assert abs(loc[i] * out[i, 1] - 1.0 / SD ** 2 - g_o[0]) <= 1e-12 * max(abs(g_o[0]), 1.0)
            dc = -cuts / SD ** 2                                 # prior on the constrained cut-points
            if cls[i] != Cn:
                dc[cls[i] - 1] += out[i, 3]
            if cls[i] != 1:
                dc[cls[i] - 2] -= out[i, 2]
            du = np.array([dc[k:].sum() * (1.0 if k == 0 else np.exp(u[k])) for k in range(Cn - 1)])
            assert np.max(np.abs(du - g_o[1:])) <= 1e-12 * max(np.max(np.abs(g_o[1:])), 1.0), (Cn, loc[i], cls[i])


def test_rows_categorical_logit(host_lib):
    """categorical_logit_row (not used by a kernel yet) against the oracle's categorical model on one-row data with
    K = 1, x = 1: lin_c = alpha_c + beta_c; d/dalpha_c = d/dbeta_c (before the priors) = the per-class weight."""
    rng = np.random.default_rng(8)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for Cn in (2, 4, 9):
        n = 60
        alpha, beta = rng.normal(0, 1.5, (n, Cn)), rng.normal(0, 1.5, (n, Cn))
        alpha[0], beta[0] = 0.0, 0.0
        alpha[1, 0] = 40.0                                       # one class dominating: exp underflow in the others
        y = rng.integers(1, Cn + 1, n).astype(np.int32)
        lin = np.ascontiguousarray(alpha + beta)
        lp = np.empty(n)
        host_lib.categorical_logit_rows(n, Cn, y.ctypes.data_as(C.POINTER(C.c_int)), dp(lin), dp(lp))
        for i in range(n):
            po = PortOracle("categorical_logit", np.array([[1.0]]), np.array([y[i]]), n_classes=Cn)
            lp_o, g_o = po.log_prob_grad(np.concatenate([alpha[i], beta[i]]), propto=True, jacobian=False)
            prior = -0.5 * np.sum((alpha[i] / SD) ** 2) - 0.5 * np.sum((beta[i] / SD) ** 2)
            assert abs(lp[i] + prior - lp_o) <= 1e-13 * max(abs(lp_o), 1.0), (Cn, i)
            assert np.max(np.abs(lin[i] - alpha[i] / SD ** 2 - g_o[:Cn])) <= 1e-13 * max(np.max(np.abs(g_o)), 1.0)
            assert np.max(np.abs(lin[i] - beta[i] / SD ** 2 - g_o[Cn:])) <= 1e-13 * max(np.max(np.abs(g_o)), 1.0)


# ---- the branch-free link step of glm_multi_kernel (link_bf<>, fm_exp / fm_log1p in glm_link.cuh) ----
def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_fast_exp_and_log1p_against_libm(host_lib):
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-700, 700, 200_000), rng.uniform(-25, 25, 200_000), rng.uniform(-1e-3, 1e-3, 1000),
                        [0.0, -0.0, 700.0, -700.0, 1e-300, 20.0, -20.0, 21.0]])
    out = np.empty_like(x)
    host_lib.fm_exp_rows(x.size, dp(x), dp(out))
    assert _ulps(out, np.exp(x)).max() <= 1.0                       # the fast path of CUDA's exp: < 1 ulp
    u = np.concatenate([np.exp(rng.uniform(-50, 21, 300_000)), rng.uniform(0, 3, 100_000), [0.0, 1e-300, 1.0, 0.41, 0.42,
                        np.sqrt(2) - 1, 4.85e8]])
    l1p, rw = np.empty_like(u), np.empty_like(u)
    host_lib.fm_log1p_rows(u.size, dp(u), dp(l1p), dp(rw))
    ref = np.log1p(u)
    ok = ref > 0
    assert _ulps(l1p[ok], ref[ok]).max() <= 2.0                     # fdlibm's algorithm: < 1 ulp, + the reciprocal's
    assert np.all(l1p[~ok] == 0.0)
    assert np.max(np.abs(rw * (1.0 + u) - 1.0)) < 4e-16


@pytest.mark.parametrize("fam", [0, 1, 2, 3])
def test_branch_free_link_equals_link(host_lib, fam):
    """link_bf<> against link<> row by row: random etas, the bernoulli cut-offs from both sides, huge |eta|, infinities
    and NaN (which branch is taken and what propagates must be the same), both outcomes."""
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng = np.random.default_rng(6)
    eta = np.concatenate([rng.normal(0, 3, 100_000), rng.uniform(-30, 30, 100_000),
                          [20.0, -20.0, np.nextafter(20.0, 30), np.nextafter(20.0, 0), np.nextafter(-20.0, -30),
                           np.nextafter(-20.0, 0), 0.0, 50.0, -50.0, 699.0, -699.0, 705.0, -705.0, 745.0, -745.0, 800.0,
                           -800.0, 1e6, -1e6, np.inf, -np.inf, np.nan]])
    aux = np.zeros(eta.size)
    if fam == 3:                                              # binomial_logit: population sizes incl. 0, y in [0, n]
        aux = rng.integers(0, 40, eta.size).astype(float)
        y = np.floor(rng.uniform(0, 1, eta.size) * (aux + 1))
    elif fam == 1:
        y = rng.poisson(3.0, eta.size).astype(float)
    elif fam == 0:
        y = rng.integers(0, 2, eta.size).astype(float)
    else:
        y = rng.normal(0, 2, eta.size)
    n = eta.size
    lp, r, lpb, rb = (np.empty(n) for _ in range(4))
    assert host_lib.link_pair_rows(fam, n, dp(eta), dp(y), C.c_double(0.7), dp(lp), dp(r), dp(lpb), dp(rb), dp(aux)) == 0
    with np.errstate(invalid="ignore", over="ignore"):
        for a, b in ((lp, lpb), (r, rb)):
            assert np.array_equal(np.isnan(a), np.isnan(b))
            assert np.array_equal(np.isinf(a), np.isinf(b)) and np.array_equal(a[np.isinf(a)], b[np.isinf(b)])
            fin = np.isfinite(a)
            tiny = fin & (np.abs(a) < 1e-290)                 # exp(-t) beyond t = 700: below 1e-304 in both
            assert np.all(np.abs(b[tiny]) < 1e-290)
            big = fin & ~tiny
            # poisson: y eta - exp(eta) cancels; measure against the larger term
            scale = np.maximum(np.abs(a[big]), np.abs(y[big] * eta[big]) if fam == 1 else 0.0)
            tol = 1e-15 * (8 if fam == 1 else 4)
            if fam == 3:
                # r = y - n inv_logit(eta) cancels: measure against y; the reference forms inv_logit as exp(log_inv_logit),
                # whose own rounding error grows like |eta| ulp -- the branch-free form (a reciprocal) is the tighter one
                scale = np.maximum(scale, y[big])
                tol = 4e-15 + 2.5e-16 * np.abs(eta[big])
            assert np.all(np.abs(a[big] - b[big]) / scale < tol)
