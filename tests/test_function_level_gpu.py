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
