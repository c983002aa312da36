"""GPU suite for the function-level binding: b200glm_glm_lpmf through the C ABI and the
stan::math overloads of stan_b200/cpp/b200/glm_functions.hpp (as one node of a reverse-mode tape) against the
bare reference densities -- golden vectors generated from the reference, and live against oracle/_ref."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err, rel_err_vec, unhex
from stan_b200 import GLMModel, make_glm_data, stan_service

pytestmark = pytest.mark.gpu
TOL = 1e-10


def load_cases():
    with open(os.path.join(ROOT, "tests", "golden", "glm_function_golden.json")) as f:
        g = json.load(f)
    out = []
    for c in g["cases"]:
        c["X"] = unhex(c["X"]).reshape((c["N"], c["K"]), order="F")
        c["y"] = np.array(c["y"], dtype=np.float64 if c["family"] == "normal_id" else np.int32)
        c["group"] = None if c["group"] is None else np.array(c["group"], dtype=np.int32)
        out.append(c)
    return out


CASES = load_cases()


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_c_abi_function_matches_reference_golden(c):
    m = GLMModel(c["family"], c["X"], c["y"], c["group"], c["G"])
    a, b = unhex(c["alpha"]), unhex(c["beta"])
    for e in c["evals"]:
        lp, da, db, ds = m.glm_lpmf(a, b, c["sigma"], e["propto"], e["operands_are_var"], e["sigma_is_var"])
        ref = float.fromhex(e["lp"])
        assert rel_err(lp, ref) < TOL, (c["name"], e, lp, ref)
        if e["operands_are_var"]:
            assert rel_err_vec(da, unhex(e["d_alpha"])) < TOL and rel_err_vec(db, unhex(e["d_beta"])) < TOL
            if c["family"] == "normal_id" and e["sigma_is_var"]:
                assert rel_err(ds, float.fromhex(e["d_sigma"])) < TOL
    m.close()


@pytest.mark.skipif(not stan_service.available(), reason="libb200stan.so not built")
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_stan_math_overloads_on_the_tape(c):
    """f = scale * glm(...) + 0.5 sum(beta^2) through stan::math::*_glm_lp*f(d.y(), d.x(), ...): value and adjoints."""
    fm = stan_service.FuncGLM(c["family"], c["X"], c["y"], c["group"], c["G"])
    a, b = unhex(c["alpha"]), unhex(c["beta"])
    scale = -2.5
    for e in c["evals"]:
        f, da, db, ds = fm.eval(a, b, c["sigma"], e["propto"], e["operands_are_var"], e["sigma_is_var"], scale)
        assert rel_err(f, scale * float.fromhex(e["lp"]) + 0.5 * float(b @ b)) < TOL, (c["name"], e)
        if e["operands_are_var"]:
            assert rel_err_vec(da, scale * unhex(e["d_alpha"])) < TOL
            assert rel_err_vec(db, scale * unhex(e["d_beta"]) + b) < TOL
            if c["family"] == "normal_id" and e["sigma_is_var"]:
                assert rel_err(ds, scale * float.fromhex(e["d_sigma"])) < TOL
    fm.close()


def test_function_level_live_and_errors():
    from oracle.oracle import RefOracle
    import stan_b200
    if not RefOracle.available():
        pytest.skip("oracle/_ref not present")
    d = make_glm_data("poisson_log", 40_000, 50, 1000)        # config 4's K and G
    m = GLMModel("poisson_log", d["X"], d["y"], d["group"], 1000)
    rng = np.random.default_rng(2)
    a, b = 0.3 * rng.standard_normal(1000), 0.05 * rng.standard_normal(50)
    lp, da, db, _ = m.glm_lpmf(a, b)
    lp_r, da_r, db_r, _ = RefOracle.glm_function("poisson_log", d["X"], d["y"], a, b, 1.0, d["group"], 1000)
    assert rel_err(lp, lp_r) < TOL and rel_err_vec(da, da_r) < TOL and rel_err_vec(db, db_r) < TOL
    with pytest.raises(stan_b200.InvalidArgument):
        m.glm_lpmf(a[:5], b)
    m.close()
    d = make_glm_data("normal_id", 1000, 3)
    m = GLMModel("normal_id", d["X"], d["y"])
    with pytest.raises(stan_b200.DomainError):                # check_positive_finite(sigma), normal_id_glm_lpdf.hpp:93
        m.glm_lpmf(0.0, np.zeros(3), sigma=-1.0)
    m.close()


# ---------------------------------------------------------------------------------------------
# Per-row operands: an N-vector intercept (every family) and an N-vector scale (normal_id) --
# the vector forms of the reference's overloads (SM/opencl/prim/normal_id_glm_lpdf.hpp:68-84).
# Checked against the CPU checker WITHOUT new oracle code, through two identities:
#   * an intercept vector a is one more column of X with weight 1:  glm(X, a, beta) == glm([X a], 0, [beta 1]),
#     and a one-hot column e_j reads out row j's partial:           d/d weight of e_j == d/d a_j;
#   * a per-row scale rescales the data:  normal(y | X beta + a, s_i) == normal(y / s | (X / s) beta + a / s, 1) - sum log s_i.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fam", ["bernoulli_logit", "poisson_log", "normal_id", "binomial_logit", "neg_binomial_2_log"])
def test_per_row_intercept_matches_checker(fam):
    from oracle.oracle import RefOracle, PortOracle
    from stan_b200 import GLMModel, make_glm_data
    N, K = 3_003, 7
    d = make_glm_data(fam, N, K)
    kw = {"trials": d["trials"]} if fam == "binomial_logit" else {}
    rng = np.random.default_rng(8)
    a, beta, sig = 0.3 * rng.standard_normal(N), 0.2 * rng.standard_normal(K), 1.7
    m = GLMModel(fam, d["X"], d["y"], **kw)
    lp, da, db, ds = m.glm_lpmf_rows(beta, alpha_rows=a, sigma=sig)
    m.close()
    rows = [0, 1, 31, 32, 1500, N - 1]
    onehot = np.zeros((N, len(rows)))
    onehot[rows, np.arange(len(rows))] = 1.0
    Xa = np.asfortranarray(np.column_stack([d["X"], a, onehot]))
    ba = np.concatenate([beta, [1.0], np.zeros(len(rows))])
    if RefOracle.available():
        lp_r, _, db_r, ds_r = RefOracle.glm_function(fam, Xa, d["y"], 0.0, ba, sig, **kw)
    else:
        pytest.skip("needs oracle/_ref for the bare density")
    assert rel_err(lp, lp_r) < 1e-10
    assert rel_err_vec(db, db_r[:K]) < 1e-10
    assert rel_err_vec(da[rows], db_r[K + 1:]) < 1e-10                 # per-row partials, read out by the one-hot columns
    assert rel_err(float(a @ da), db_r[K]) < 1e-9                      # and their a-weighted sum
    if fam in ("normal_id", "neg_binomial_2_log"):
        assert rel_err(ds, ds_r) < 1e-10


def test_per_row_scale_matches_rescaled_model():
    from oracle.oracle import RefOracle
    from stan_b200 import GLMModel, make_glm_data
    if not RefOracle.available():
        pytest.skip("needs oracle/_ref for the bare density")
    N, K = 4_001, 5
    d = make_glm_data("normal_id", N, K)
    rng = np.random.default_rng(9)
    s = np.exp(0.4 * rng.standard_normal(N))
    a, alpha0, beta = 0.3 * rng.standard_normal(N), 0.25, 0.2 * rng.standard_normal(K)
    m = GLMModel("normal_id", d["X"], d["y"])
    for arows in (None, a):
        off = np.full(N, alpha0) if arows is None else arows
        lp, da, db, dsr = m.glm_lpmf_rows(beta, alpha=alpha0, alpha_rows=arows, sigma_rows=s, propto=False)
        Xs = np.asfortranarray(np.column_stack([d["X"] / s[:, None], off / s]))
        lp_r, _, db_r, _ = RefOracle.glm_function("normal_id", Xs, d["y"] / s, 0.0, np.concatenate([beta, [1.0]]), 1.0,
                                                  propto=False)
        assert rel_err(lp, lp_r - np.log(s).sum()) < 1e-10
        assert rel_err_vec(db, db_r[:K]) < 1e-10
        z = (d["y"] - d["X"] @ beta - off) / s
        assert rel_err_vec(dsr, (z * z - 1.0) / s) < 1e-10             # normal_id_glm_lpdf.hpp:181-183 per row
        if arows is not None:
            assert rel_err_vec(da, z / s) < 1e-10                      # mu_derivative
        else:
            assert rel_err(da, float((z / s).sum())) < 1e-10
    # propto with sigma a constant drops -sum log sigma_i; all-constant operands give 0
    lp_p, *_ = m.glm_lpmf_rows(beta, alpha=alpha0, sigma_rows=s, propto=True, sigma_is_var=False)
    lp_f, *_ = m.glm_lpmf_rows(beta, alpha=alpha0, sigma_rows=s, propto=True, sigma_is_var=True)
    assert rel_err(lp_f - lp_p, -np.log(s).sum()) < 1e-10
    assert m.glm_lpmf_rows(beta, alpha=alpha0, sigma_rows=s, propto=True, operands_are_var=False)[0] == 0.0
    from stan_b200.model import DomainError, InvalidArgument
    bad = s.copy()
    bad[17] = 0.0
    with pytest.raises(DomainError):
        m.glm_lpmf_rows(beta, sigma_rows=bad)
    with pytest.raises(InvalidArgument):
        m.glm_lpmf_rows(beta, sigma_rows=s[:-1])
    m.close()


@pytest.mark.skipif(not stan_service.available(), reason="libb200stan.so not built")
@pytest.mark.parametrize("fam,vec_sigma", [("bernoulli_logit", False), ("neg_binomial_2_log", False),
                                           ("normal_id", False), ("normal_id", True)])
def test_per_row_overloads_on_the_tape(fam, vec_sigma):
    """stan::math::<family>_glm_lp*f(d.y(), d.x(), b200::by_row(alpha), beta [, sigma vector]) with every operand a
    var, as ONE node of a larger reverse-mode tape: adjoints == scale x the C entry's partials (+ beta for the
    0.5 sum beta^2 term)."""
    N, K = 2_050, 6
    d = make_glm_data(fam, N, K)
    rng = np.random.default_rng(12)
    a, beta = 0.3 * rng.standard_normal(N), 0.2 * rng.standard_normal(K)
    s = np.exp(0.3 * rng.standard_normal(N)) if vec_sigma else None
    sig, scale = 1.4, 0.7
    fg = stan_service.FuncGLM(fam, d["X"], d["y"])
    f, da, db, ds = stan_service.func_eval_rows(fg, a, beta, sigma_rows=s, sigma=sig, propto=True, scale=scale)
    fg.close()
    m = GLMModel(fam, d["X"], d["y"])
    lp, da_c, db_c, ds_c = m.glm_lpmf_rows(beta, alpha_rows=a, sigma_rows=s, sigma=sig, propto=True)
    m.close()
    assert rel_err(f, scale * lp + 0.5 * float(beta @ beta)) < 1e-12
    assert rel_err_vec(da, scale * da_c) < 1e-12
    assert rel_err_vec(db, scale * db_c + beta) < 1e-12
    if vec_sigma:
        assert rel_err_vec(ds, scale * ds_c) < 1e-12
    elif fam in ("normal_id", "neg_binomial_2_log"):
        assert rel_err(ds, scale * ds_c) < 1e-12
