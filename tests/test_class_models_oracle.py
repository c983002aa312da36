"""CPU suite: the oracle for the reference's last two GLMs, ordered_logistic_glm_lpmf and categorical_logit_glm_lpmf
(SURVEY 8f row 3).  These are NOT built on the device yet; per the scope order the oracle
comes first: the plain-C port is pinned here to the compiled reference -- committed goldens
(tests/golden/glm_class_models_golden.json, `make_golden.py classes`) and live against oracle/_ref -- so the
CUDA path has its checker waiting.
"""
import numpy as np
import pytest

from conftest import CLASS_GOLDEN_NAMES, rel_err, rel_err_vec, unhex
from oracle.oracle import OracleError, PortOracle, RefOracle
from stan_b200.synth import make_glm_data

TOL = 1e-12


@pytest.mark.parametrize("name", CLASS_GOLDEN_NAMES)
def test_port_matches_golden(golden_classes, name):
    c = golden_classes[name]
    po = PortOracle(c["family"], c["X"], c["y"], n_classes=c["n_classes"])
    assert po.P == len(c["evals"][0]["theta"])
    for e in c["evals"]:
        th = unhex(e["theta"])
        for key, ref in e["lp_grad"].items():
            lp, g = po.log_prob_grad(th, int(key[0]), int(key[1]))
            assert rel_err(lp, float.fromhex(ref["lp"])) < TOL, (name, key)
            assert rel_err_vec(g, unhex(ref["grad"])) < TOL, (name, key)
        for key, ref in e["lp_double"].items():
            assert rel_err(po.log_prob(th, int(key[0]), int(key[1])), float.fromhex(ref)) < TOL, (name, key)
    lf = c["leapfrog"]
    q, p, g, V = po.leapfrog(lf["eps"], unhex(lf["inv_metric"]), unhex(lf["q0"]), unhex(lf["p0"]), unhex(lf["g0"]),
                             float.fromhex(lf["V0"]))
    assert rel_err_vec(q, unhex(lf["q1"])) < TOL and rel_err_vec(p, unhex(lf["p1"])) < TOL
    assert rel_err_vec(g, unhex(lf["g1"])) < TOL and rel_err(V, float.fromhex(lf["V1"])) < TOL


@pytest.mark.skipif(not RefOracle.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("fam,N,K,C", [("ordered_logistic", 2000, 20, 5), ("ordered_logistic", 999, 3, 9),
                                       ("categorical_logit", 1500, 12, 4), ("categorical_logit", 700, 30, 7)])
def test_port_matches_reference_live(fam, N, K, C):
    d = make_glm_data(fam, N, K, n_classes=C)
    po = PortOracle(fam, d["X"], d["y"], n_classes=C)
    ro = RefOracle(fam, d["X"], d["y"], n_classes=C)
    rng = np.random.default_rng(17)
    for sc in (0.0, 0.2, 1.0):
        th = sc * rng.standard_normal(po.P)
        for propto in (1, 0):
            lp1, g1 = po.log_prob_grad(th, propto, 1)
            lp2, g2 = ro.log_prob_grad(th, propto, 1)
            assert rel_err(lp1, lp2) < TOL and rel_err_vec(g1, g2) < TOL
        assert rel_err(po.log_prob(th, 0, 1), ro.log_prob(th, 0, 1)) < TOL


def test_error_behaviour_and_quirks():
    d = make_glm_data("ordered_logistic", 30, 2, n_classes=3)
    y = d["y"].copy()
    y[4] = 4                                              # check_bounded(y, 1, N_classes), ordered...:84
    with pytest.raises(OracleError):
        PortOracle("ordered_logistic", d["X"], y, n_classes=3).log_prob_grad(np.zeros(4))
    po = PortOracle("ordered_logistic", d["X"], d["y"], n_classes=3)
    with pytest.raises(OracleError):                      # exp(u) underflows: cut-points not strictly increasing, :85
        po.log_prob_grad(np.array([0.0, 0.0, 0.3, -800.0]))
    # one class: no cut-points, size_zero(cuts) => the likelihood is 0, priors remain (:93-95)
    po = PortOracle("ordered_logistic", d["X"], np.ones(30, np.int32), n_classes=1)
    lp, g = po.log_prob_grad(np.array([0.5, -0.5]))
    assert rel_err(lp, -0.5 * 2 * (0.5 / 2.5) ** 2) < 1e-15
    d = make_glm_data("categorical_logit", 30, 2, n_classes=3)
    y = d["y"].copy()
    y[0] = 0
    with pytest.raises(OracleError):                      # categorical outcome out of support, categorical...:74
        PortOracle("categorical_logit", d["X"], y, n_classes=3).log_prob_grad(np.zeros(9))
    # N_classes == 1 returns 0 BEFORE the bounds check (:70-72): a bad y is not an error there
    po = PortOracle("categorical_logit", d["X"], y, n_classes=1)
    lp, g = po.log_prob_grad(np.zeros(3))
    assert lp == 0.0 and not g.any()
