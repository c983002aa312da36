"""CPU suite: pins the plain-C oracle (oracle/glm_oracle.c) to the compiled reference.

 * against the committed golden vectors generated from the reference (tests/golden/make_golden.py)
 * live against oracle/_ref/libref_oracle_*.so when it is present (this container)
Tolerance 1e-12 relative: the port sums with compensated accumulation, the reference with Eigen's
vectorised reductions; both are fp64, agreement is ~1e-15..1e-13.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_NAMES, ROOT, rel_err, rel_err_vec, unhex
from oracle.oracle import OracleError, PortOracle, RefOracle
from stan_b200.synth import make_glm_data, theta_points

TOL = 1e-12


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_port_matches_golden(golden, name):
    c = golden[name]
    po = PortOracle(c["family"], c["X"], c["y"], c["group"], c["G"])
    for e in c["evals"]:
        th = unhex(e["theta"])
        for key, ref in e["lp_grad"].items():
            propto, jac = int(key[0]), int(key[1])
            lp, g = po.log_prob_grad(th, propto, jac)
            assert rel_err(lp, float.fromhex(ref["lp"])) < TOL, (name, key)
            assert rel_err_vec(g, unhex(ref["grad"])) < TOL, (name, key)
        for key, ref in e["lp_double"].items():
            propto, jac = int(key[0]), int(key[1])
            assert rel_err(po.log_prob(th, propto, jac), float.fromhex(ref)) < TOL, (name, key)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_port_leapfrog_matches_golden(golden, name):
    c = golden[name]
    lf = c["leapfrog"]
    po = PortOracle(c["family"], c["X"], c["y"], c["group"], c["G"])
    q, p, g, V = po.leapfrog(lf["eps"], unhex(lf["inv_metric"]), unhex(lf["q0"]), unhex(lf["p0"]), unhex(lf["g0"]),
                             float.fromhex(lf["V0"]))
    assert rel_err_vec(q, unhex(lf["q1"])) < TOL
    assert rel_err_vec(p, unhex(lf["p1"])) < TOL
    assert rel_err_vec(g, unhex(lf["g1"])) < TOL
    assert rel_err(V, float.fromhex(lf["V1"])) < TOL


def test_bernoulli_cutoff_quirk_is_kept():
    """|ytheta| > 20 branches incl. the sign quirk of bernoulli_logit_glm_lpmf.hpp:137-142."""
    X = np.array([[30.0], [-30.0], [30.0], [-30.0], [0.5]])
    y = np.array([1, 1, 0, 0, 1], dtype=np.int32)
    po = PortOracle("bernoulli_logit", X, y)
    lp, g = po.log_prob_grad(np.array([0.0, 1.0]))
    e = np.exp(-30.0)
    t5 = 0.5
    want = -e + (-30.0) + (-30.0) + (-e) - np.log1p(np.exp(-t5)) - 0.5 * (1.0 / 2.5) ** 2
    assert rel_err(lp, want) < 1e-14
    r = np.array([-e, 1.0, -1.0, -e, np.exp(-t5) / (1 + np.exp(-t5))])
    assert rel_err_vec(g, [r.sum(), (X[:, 0] * r).sum() - 1.0 / 2.5 ** 2]) < 1e-14


def test_error_behaviour():
    d = make_glm_data("bernoulli_logit", 10, 2)
    y = d["y"].copy()
    y[3] = 2
    po = PortOracle("bernoulli_logit", d["X"], y)
    with pytest.raises(OracleError) as ei:
        po.log_prob_grad(np.zeros(3))
    assert ei.value.code == 1
    d = make_glm_data("poisson_log", 10, 2)
    y = d["y"].copy()
    y[0] = -1
    with pytest.raises(OracleError):
        PortOracle("poisson_log", d["X"], y).log_prob_grad(np.zeros(3))
    po = PortOracle("poisson_log", d["X"], d["y"])
    with pytest.raises(OracleError):
        po.log_prob_grad(np.array([np.inf, 0, 0]))
    with pytest.raises(OracleError):   # exp overflow => non-finite density
        po.log_prob_grad(np.array([800.0, 0, 0]))


def test_empty_data():
    """size_zero(y) => the likelihood contributes 0 (bernoulli :80-82); priors remain."""
    X = np.zeros((0, 3))
    for fam, y in (("bernoulli_logit", np.zeros(0, np.int32)), ("normal_id", np.zeros(0))):
        po = PortOracle(fam, X, y)
        th = 0.1 * np.arange(1, po.P + 1)
        lp, g = po.log_prob_grad(th)
        assert np.isfinite(lp) and np.all(np.isfinite(g))


@pytest.mark.skipif(not RefOracle.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("fam,N,K,G", [("bernoulli_logit", 2000, 20, 0), ("poisson_log", 1500, 12, 0),
                                       ("normal_id", 1800, 16, 0), ("poisson_log", 3000, 8, 37),
                                       ("bernoulli_logit", 10000, 20, 0)])
def test_port_matches_reference_live(fam, N, K, G):
    d = make_glm_data(fam, N, K, G)
    po = PortOracle(fam, d["X"], d["y"], d["group"], G)
    ro = RefOracle(fam, d["X"], d["y"], d["group"], G)
    for th in theta_points(po.P, n_random=2, scale=0.2):
        for propto in (1, 0):
            lp1, g1 = po.log_prob_grad(th, propto, 1)
            lp2, g2 = ro.log_prob_grad(th, propto, 1)
            assert rel_err(lp1, lp2) < TOL and rel_err_vec(g1, g2) < TOL
        assert rel_err(po.log_prob(th, 0, 1), ro.log_prob(th, 0, 1)) < TOL
    lp_g, g_g = ro.gradient(th)                   # stan::model::gradient == log_prob_grad<true,true>
    assert lp_g == lp2 or rel_err(lp_g, ro.log_prob_grad(th)[0]) < 1e-15


@pytest.mark.skipif(not RefOracle.available(), reason="oracle/_ref not built")
def test_reference_ess_golden():
    """stan::analyze::ess wiring check on an AR(1) chain (the ESS used for ESS/s)."""
    rng = np.random.default_rng(3)
    x = np.zeros((1000, 4))
    for c in range(4):
        for i in range(1, 1000):
            x[i, c] = 0.5 * x[i - 1, c] + rng.standard_normal()
    ess = RefOracle.ess(x)
    assert 800 < ess < 2200   # theory: 4000 * (1-0.5)/(1+0.5) = 1333


def test_port_oracle_consistent_with_function_level_goldens():
    """tests/golden/glm_function_golden.json holds the BARE reference densities (no priors).  The plain-C
    port evaluates the whole model; model lp (propto, no Jacobian) minus the closed-form normal priors must
    reproduce them, and likewise for the beta gradient."""
    import json
    from oracle.oracle import PortOracle
    with open(os.path.join(ROOT, "tests", "golden", "glm_function_golden.json")) as f:
        cases = json.load(f)["cases"]
    unhex = lambda a: np.array([float.fromhex(v) for v in a])
    for c in cases:
        N, K, G, fam = c["N"], c["K"], c["G"], c["family"]
        X = unhex(c["X"]).reshape((N, K), order="F")
        y = np.array(c["y"], dtype=np.float64 if fam == "normal_id" else np.int32)
        grp = None if c["group"] is None else np.array(c["group"], dtype=np.int32)
        po = PortOracle(fam, X, y, grp, G)
        a, b, sigma = unhex(c["alpha"]), unhex(c["beta"]), c["sigma"]
        th = np.concatenate([[0.1, -0.2], a, b] if G else [a, b])
        if fam == "normal_id":
            th = np.concatenate([th, [np.log(sigma)]])
        lp, g = po.log_prob_grad(th, True, False)
        if G:
            mu, sa = th[0], np.exp(th[1])
            pri = -0.5 * (mu / 2.5) ** 2 - 0.5 * sa ** 2 - 0.5 * np.sum(((a - mu) / sa) ** 2) - G * np.log(sa)
        else:
            pri = -0.5 * (a[0] / 2.5) ** 2
        pri += -0.5 * np.sum((b / 2.5) ** 2)
        if fam == "normal_id":
            pri += -0.5 * ((sigma - 1.0) / 2.0) ** 2
        e = [e for e in c["evals"] if e["propto"] == 1 and e["operands_are_var"] == 1
             and e["sigma_is_var"] == (1 if fam == "normal_id" else 0)][0]
        ref_lp, ref_db = float.fromhex(e["lp"]), unhex(e["d_beta"])
        assert abs((lp - pri) - ref_lp) <= 1e-11 * abs(ref_lp), c["name"]
        ob = 2 + G if G else 1
        assert np.max(np.abs(g[ob:ob + K] + b / 2.5 ** 2 - ref_db)) <= 1e-11 * np.max(np.abs(ref_db)), c["name"]


def load_normal_glm_fixture():
    """The reference's own normal_id_glm posterior fixture (tests/golden/make_normal_glm_fixture.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "normal_glm_data.json")) as f:
        d = json.load(f)
    with open(os.path.join(ROOT, "tests", "golden", "normal_glm_expected.json")) as f:
        exp = json.load(f)
    X, y = np.array(d["X"], dtype=np.float64), np.array(d["Y"], dtype=np.float64)
    assert X.shape == (d["N"], d["K"])
    return X, y, exp


def check_against_normal_glm_answers(draws, means_X, exp):
    """draws: (chains, iterations, P) constrained draws in OUR order [Intercept, b_1..b_K, sigma].  normal_glm.stan
    computes the centred Xc but hands the UN-centred X to normal_id_glm_lupdf (line 27), so `Intercept` is the
    intercept on the raw predictors and b_Intercept = Intercept - dot(means_X, b) is just a derived column.
    Compares with the reference's published posterior means / SDs of b, Intercept, sigma and b_Intercept at the
    reference's own bars (normal_glm_test.cpp:130-135)."""
    K = len(means_X)
    flat = draws.reshape(-1, draws.shape[2])
    cols = {f"b.{k + 1}": flat[:, 1 + k] for k in range(K)}
    cols["Intercept"] = flat[:, 0]
    cols["sigma"] = flat[:, K + 1]
    cols["b_Intercept"] = flat[:, 0] - flat[:, 1:1 + K] @ means_X
    worst = 0.0
    for name, m_ref, s_ref in zip(exp["names"], exp["mean"], exp["sd"]):
        if name in ("lp_approx__", "lp__"):
            continue                       # the reference's test skips these two columns as well
        x = cols[name]
        assert abs(x.mean() - m_ref) < exp["bars"]["mean_abs"], (name, x.mean(), m_ref)
        assert abs(x.std(ddof=1) - s_ref) < exp["bars"]["sd_abs"], (name, x.std(ddof=1), s_ref)
        worst = max(worst, abs(x.mean() - m_ref) / s_ref)
    # far tighter than the reference's bar: every mean within a quarter of a posterior SD of the published value
    assert worst < 0.25, worst
    return worst


def test_compiled_reference_reproduces_the_references_posterior_fixture():
    """Pins the oracle model (priors, Jacobian, parameterisation) to the one known-answer posterior the
    reference holds for a GLM: normal_glm.stan on normal_glm_test.json, answers from pathfinder/util.hpp."""
    from oracle.oracle import RefOracle
    if not RefOracle.available():
        pytest.skip("oracle/_ref not built")
    X, y, exp = load_normal_glm_fixture()
    means = X.mean(axis=0)
    ro = RefOracle("normal_id", X, y, **{k: v for k, v in exp["priors"].items() if k != "source"})
    res = ro.nuts(num_chains=4, seed=2026, num_warmup=500, num_samples=500, delta=0.8, num_threads=4)
    check_against_normal_glm_answers(res["draws"][:, :, 7:], means, exp)
