"""CPU suite: the device path's arithmetic end to end, compiled for the host.

The kernels' row arithmetic (stan_b200/csrc/glm_link.cuh) and the model epilogue finish() (glm_model.cuh: priors,
lb_constrain Jacobians, gradient assembly incl. the hierarchical mu_a / sigma_a entries, domain status, second
half of the leapfrog step) are the SAME SOURCE the GPU executes, built here with g++ as a one-thread CTA
(tests/host/link_host.cpp).  The likelihood sums finish() consumes are formed on the host from the link step's
per-row outputs exactly as the kernels lay them out (lik[P + 2]); the result is compared with the CPU oracle --
stan::model::log_prob_grad semantics (var), Model::log_prob<propto, jacobian>(double) semantics, and one
expl_leapfrog step.  What this leaves to the -m gpu suite is the parallel machinery: TMA staging, the warp /
CTA / cross-CTA reductions, the peer exchange.
"""
import ctypes as C

import numpy as np
import pytest
from scipy.special import gammaln

from conftest import rel_err, rel_err_vec
from oracle.oracle import PortOracle
from stan_b200.synth import make_glm_data
from test_link_math_host import host_lib, link_rows  # noqa: F401  (fixture + helper)

FAM = {"bernoulli_logit": 0, "poisson_log": 1, "normal_id": 2, "binomial_logit": 3, "neg_binomial_2_log": 4}
PRIORS = dict(prior_alpha_sd=2.5, prior_beta_sd=1.7, prior_sigma_loc=0.8, prior_sigma_scale=2.2, prior_sigma_a_scale=1.3)
TOL = 1e-12


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class HostDevicePath:
    """The device path of one model, run on the host: link rows -> likelihood sums -> finish()."""

    def __init__(self, L, family, d, G):
        self.L, self.family, self.fam, self.d, self.G = L, family, FAM[family], d, G
        self.X, self.y = np.asarray(d["X"]), np.asarray(d["y"], dtype=np.float64)
        self.N, self.K = self.X.shape
        self.has_scale = self.fam in (2, 4)
        self.off_beta = 2 + G if G else 1
        self.P = self.off_beta + self.K + (1 if self.has_scale else 0)
        self.aux = np.asarray(d["trials"], dtype=np.float64) if "trials" in d else None
        if self.fam in (1, 4):
            self.const = float(np.sum(gammaln(self.y + 1.0)))
        elif self.fam == 3:
            n, k = self.aux, np.minimum(self.y, self.aux - self.y)
            self.const = -float(np.sum(np.where(k == 0, 0.0, gammaln(n + 1) - gammaln(k + 1) - gammaln(n + 1 - k))))
        else:
            self.const = 0.0
        L.finish_host.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double)] + [C.POINTER(C.c_double)] * 5
        L.finish_host.restype = None

    def lik_sums(self, theta):
        """lik[P + 2] as cross_cta_reduce_and_finish / group_reduce_kernel leave it."""
        beta = theta[self.off_beta:self.off_beta + self.K]
        off = theta[2:2 + self.G][self.d["group"] - 1] if self.G else theta[0]
        eta = np.ascontiguousarray(self.X @ beta + off)
        scale = float(np.exp(theta[-1])) if self.has_scale else 1.0
        lp, r, x = link_rows(self.L, self.fam, eta, self.y, self.aux, scale)
        lik = np.zeros(self.P + 2)
        if self.G:
            np.add.at(lik, 1 + self.d["group"], r)        # lik[2 + (group - 1)]
        else:
            lik[0] = r.sum()
        lik[self.off_beta:self.off_beta + self.K] = self.X.T @ r
        lik[self.P], lik[self.P + 1] = lp.sum(), x.sum()
        return lik

    def finish(self, theta, propto, jacobian, is_var, mode=0, eps=0.0, st_in=None):
        ic = (C.c_int * 11)(self.fam, self.K, self.G, self.P, self.off_beta, int(propto), int(jacobian), int(is_var), 0, 0, mode)
        dc = (C.c_double * 8)(float(self.N), self.const, PRIORS["prior_alpha_sd"], PRIORS["prior_beta_sd"],
                              PRIORS["prior_sigma_loc"], PRIORS["prior_sigma_scale"], PRIORS["prior_sigma_a_scale"], eps)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        lik, res = self.lik_sums(theta), np.zeros(self.P + 2)
        st_in = np.zeros(3 * self.P + 1) if st_in is None else np.ascontiguousarray(st_in)
        st_out = np.zeros(3 * self.P + 1)
        self.L.finish_host(ic, dc, dp(theta), dp(lik), dp(res), dp(st_in), dp(st_out))
        return res, st_out


CASES = [("bernoulli_logit", 0), ("bernoulli_logit", 5), ("poisson_log", 0), ("poisson_log", 7), ("normal_id", 0),
         ("normal_id", 4), ("binomial_logit", 0), ("binomial_logit", 3), ("neg_binomial_2_log", 0), ("neg_binomial_2_log", 6)]


@pytest.mark.parametrize("family,G", CASES)
def test_epilogue_matches_oracle(host_lib, family, G):
    d = make_glm_data(family, 301, 5, G, seed=77 + G)
    kw = {"trials": d["trials"]} if "trials" in d else {}
    po = PortOracle(family, d["X"], d["y"], d["group"], G, **kw, **PRIORS)
    hp = HostDevicePath(host_lib, family, d, G)
    assert hp.P == po.P
    rng = np.random.default_rng(5 + G)
    for sc in (0.0, 0.2, 0.7):
        th = sc * rng.standard_normal(hp.P)
        for propto in (1, 0):
            for jac in (1, 0):
                res, _ = hp.finish(th, propto, jac, is_var=1)                 # stan::model::log_prob_grad semantics
                lp_o, g_o = po.log_prob_grad(th, propto, jac)
                assert res[hp.P + 1] == 0.0
                assert rel_err(res[0], lp_o) < TOL, (family, G, sc, propto, jac, res[0], lp_o)
                assert rel_err_vec(res[1:1 + hp.P], g_o) < TOL, (family, G, sc, propto, jac)
                res, _ = hp.finish(th, propto, jac, is_var=0)                 # Model::log_prob<propto, jacobian>(double)
                assert rel_err(res[0], po.log_prob(th, propto, jac)) < TOL, (family, G, sc, propto, jac)


@pytest.mark.parametrize("family,G", [("bernoulli_logit", 0), ("poisson_log", 7), ("normal_id", 0), ("neg_binomial_2_log", 0)])
def test_epilogue_leapfrog_tail(host_lib, family, G):
    """begin_update_p + update_q restated here (the kernels' prologue), gradient at q_new through the host device
    path, end_update_p by finish() in MODE_LEAPFROG == the oracle's expl_leapfrog step."""
    d = make_glm_data(family, 257, 4, G, seed=31)
    po = PortOracle(family, d["X"], d["y"], d["group"], G, **PRIORS)
    hp = HostDevicePath(host_lib, family, d, G)
    rng = np.random.default_rng(9)
    q0, p0 = 0.1 * rng.standard_normal(hp.P), rng.standard_normal(hp.P)
    im = np.exp(0.3 * rng.standard_normal(hp.P))
    lp0, gr0 = po.log_prob_grad(q0)
    g0, V0, eps = -gr0, -lp0, 0.013
    q_new = q0 + eps * (im * (p0 - 0.5 * eps * g0))
    _, st = hp.finish(q_new, 1, 1, 1, mode=1, eps=eps, st_in=np.concatenate([q0, p0, g0, [V0]]))
    q1, p1, g1, V1 = po.leapfrog(eps, im, q0, p0, g0, V0)
    P = hp.P
    assert rel_err_vec(st[:P], q1) < TOL and rel_err_vec(st[P:2 * P], p1) < 1e-11
    assert rel_err_vec(st[2 * P:3 * P], g1) < 1e-11 and rel_err(st[3 * P], V1) < TOL


def test_epilogue_domain_error_status(host_lib):
    """Non-finite parameter -> status ST_DOMAIN; in leapfrog mode V = +inf and g is negated (base_hamiltonian.hpp:65-69)."""
    d = make_glm_data("poisson_log", 64, 3, seed=3)
    hp = HostDevicePath(host_lib, "poisson_log", d, 0)
    th = np.array([800.0, 0.0, 0.0, 0.0])                  # exp overflow in every row
    with np.errstate(all="ignore"):
        res, _ = hp.finish(th, 1, 1, 1)
    assert res[hp.P + 1] == 1.0
    q0, p0, g0 = np.zeros(4), np.ones(4), np.array([1.0, -2.0, 3.0, -4.0])
    with np.errstate(all="ignore"):
        _, st = hp.finish(th, 1, 1, 1, mode=1, eps=0.1, st_in=np.concatenate([q0, p0, g0, [5.0]]))
    assert st[12] == np.inf and np.array_equal(st[8:12], -g0)


# ---------------------------------------------------------------------------------------------------------------
# The class models (ordered_logistic_glm, categorical_logit_glm): row arithmetic + epilogue written ahead of their
# kernels and pinned here, through the host build, to the oracle.
# ---------------------------------------------------------------------------------------------------------------
class HostClassModelPath:
    def __init__(self, L, family, d, Cn):
        self.L, self.ordered, self.Cn = L, family == "ordered_logistic", Cn
        self.X, self.y = np.asarray(d["X"]), np.ascontiguousarray(d["y"], dtype=np.int32)
        self.N, self.K = self.X.shape
        self.P = self.K + Cn - 1 if self.ordered else Cn * (1 + self.K)
        ip = C.POINTER(C.c_int)
        L.finish_class_model_host.argtypes = [ip, C.POINTER(C.c_double)] + [C.POINTER(C.c_double)] * 6
        L.finish_class_model_host.restype = None

    def lik_sums(self, theta):
        K, Cn, N = self.K, self.Cn, self.N
        lik = np.zeros(self.P + 1)
        ip = C.POINTER(C.c_int)
        if self.ordered:
            u = theta[K:]
            cuts = np.concatenate([[u[0]], u[0] + np.cumsum(np.exp(u[1:]))]) if Cn > 1 else np.zeros(0)
            loc = np.ascontiguousarray(self.X @ theta[:K])
            out = np.empty(4 * N)
            self.L.ordered_logistic_rows(N, dp(loc), self.y.ctypes.data_as(ip), Cn, dp(np.ascontiguousarray(cuts)), dp(out))
            out = out.reshape(-1, 4)
            lik[:K] = self.X.T @ out[:, 1]
            for i in range(N):                                  # ordered_logistic_glm_lpmf.hpp:202-207
                c = self.y[i]
                if c != Cn:
                    lik[K + c - 1] += out[i, 3]
                if c != 1:
                    lik[K + c - 2] -= out[i, 2]
            lik[self.P] = out[:, 0].sum()
        else:
            alpha, B = theta[:Cn], theta[Cn:].reshape(Cn, K).T    # beta[k + K c]
            lin = np.ascontiguousarray(self.X @ B + alpha)
            lp = np.empty(N)
            self.L.categorical_logit_rows(N, Cn, self.y.ctypes.data_as(ip), dp(lin), dp(lp))
            lik[:Cn] = lin.sum(axis=0)
            lik[Cn:self.P] = (self.X.T @ lin).T.ravel()
            lik[self.P] = lp.sum()
        return lik

    def finish(self, theta, propto, jacobian, is_var, mode=0, eps=0.0, st_in=None):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        ic = (C.c_int * 8)(int(self.ordered), self.K, self.Cn, self.P, int(propto), int(jacobian), int(is_var), mode)
        dc = (C.c_double * 4)(float(self.N), PRIORS["prior_alpha_sd"], PRIORS["prior_beta_sd"], eps)
        lik = self.lik_sums(theta) if self.Cn > 1 and self.N > 0 else np.zeros(self.P + 1)
        res, cuts = np.zeros(self.P + 2), np.zeros(2 * max(self.Cn, 1))
        st_in = np.zeros(3 * self.P + 1) if st_in is None else np.ascontiguousarray(st_in)
        st_out = np.zeros(3 * self.P + 1)
        self.L.finish_class_model_host(ic, dc, dp(theta), dp(lik), dp(cuts), dp(res), dp(st_in), dp(st_out))
        return res, st_out


@pytest.mark.parametrize("family,N,K,Cn", [("ordered_logistic", 301, 5, 4), ("ordered_logistic", 97, 3, 2),
                                           ("ordered_logistic", 64, 0, 3), ("ordered_logistic", 40, 2, 1),
                                           ("categorical_logit", 301, 5, 4), ("categorical_logit", 97, 2, 2),
                                           ("categorical_logit", 40, 3, 1)])
def test_class_model_epilogue_matches_oracle(host_lib, family, N, K, Cn):
    d = make_glm_data(family, N, K, n_classes=Cn, seed=91)
    po = PortOracle(family, d["X"], d["y"], n_classes=Cn, **PRIORS)
    hp = HostClassModelPath(host_lib, family, d, Cn)
    assert hp.P == po.P
    rng = np.random.default_rng(13)
    for sc in (0.0, 0.3, 0.9):
        th = sc * rng.standard_normal(hp.P)
        for propto in (1, 0):
            for jac in (1, 0):
                res, _ = hp.finish(th, propto, jac, is_var=1)
                lp_o, g_o = po.log_prob_grad(th, propto, jac)
                assert res[hp.P + 1] == 0.0
                assert rel_err(res[0], lp_o) < TOL, (family, Cn, sc, propto, jac, res[0], lp_o)
                assert rel_err_vec(res[1:1 + hp.P], g_o) < TOL, (family, Cn, sc, propto, jac)
                res, _ = hp.finish(th, propto, jac, is_var=0)
                assert rel_err(res[0], po.log_prob(th, propto, jac)) < TOL, (family, Cn, sc, propto, jac)
    # one leapfrog step
    if hp.P:
        q0, p0 = 0.1 * rng.standard_normal(hp.P), rng.standard_normal(hp.P)
        im = np.exp(0.3 * rng.standard_normal(hp.P))
        lp0, gr0 = po.log_prob_grad(q0)
        g0, V0, eps = -gr0, -lp0, 0.011
        q_new = q0 + eps * (im * (p0 - 0.5 * eps * g0))
        _, st = hp.finish(q_new, 1, 1, 1, mode=1, eps=eps, st_in=np.concatenate([q0, p0, g0, [V0]]))
        q1, p1, g1, V1 = po.leapfrog(eps, im, q0, p0, g0, V0)
        P = hp.P
        assert rel_err_vec(st[:P], q1) < TOL and rel_err_vec(st[P:2 * P], p1) < 1e-11
        assert rel_err_vec(st[2 * P:3 * P], g1) < 1e-11 and rel_err(st[3 * P], V1) < TOL


def test_class_model_epilogue_domain_status(host_lib):
    """exp(u) underflows -> cut-points not strictly increasing -> ST_DOMAIN (check_ordered, ordered...:85)."""
    d = make_glm_data("ordered_logistic", 50, 2, n_classes=3, seed=5)
    hp = HostClassModelPath(host_lib, "ordered_logistic", d, 3)
    with np.errstate(all="ignore"):
        res, _ = hp.finish(np.array([0.0, 0.0, 0.3, -800.0]), 1, 1, 1)
    assert res[hp.P + 1] == 1.0
