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
