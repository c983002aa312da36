import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def unhex(a):
    return np.array([float.fromhex(v) for v in a], dtype=np.float64)


def rel_err_vec(got, ref):
    """Gradient parity metric (SURVEY 7 'hard parts'): components are sums of +-O(1) terms that may
    cancel to ~0, so each entry is scaled by max(|g_ref,k|, ||g_ref||_inf)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    scale = np.maximum(np.abs(ref), np.max(np.abs(ref)))
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(got - ref) / scale))


def rel_err(got, ref):
    return abs(got - ref) / max(abs(ref), 1e-300) if ref != 0 else abs(got)


def _load_golden(filename):
    with open(os.path.join(ROOT, "tests", "golden", filename)) as f:
        g = json.load(f)
    cases = {}
    for c in g["cases"]:
        N, K = c["N"], c["K"]
        c["X"] = unhex(c["X"]).reshape((N, K), order="F")
        c["y"] = np.array(c["y"], dtype=np.float64 if c["family"] == "normal_id" else np.int32)
        c["group"] = None if c["group"] is None else np.array(c["group"], dtype=np.int32)
        c["trials"] = None if c.get("trials") is None else np.array(c["trials"], dtype=np.int32)
        cases[c["name"]] = c
    return cases


@pytest.fixture(scope="session")
def golden():
    return _load_golden("glm_ref_golden.json")


@pytest.fixture(scope="session")
def golden_classes():
    """ordered_logistic / categorical_logit cases (tests/golden/make_golden.py classes): oracle-side only so far."""
    return _load_golden("glm_class_models_golden.json")


@pytest.fixture(scope="session")
def golden_more():
    """binomial_logit / neg_binomial_2_log cases (tests/golden/make_golden.py more)."""
    return _load_golden("glm_more_families_golden.json")


GOLDEN_NAMES = ["bern_small", "bern_ragged", "bern_wide", "bern_groups", "pois_small", "pois_groups",
                "norm_small", "norm_ragged", "norm_groups", "bern_k1", "bern_k0", "pois_n1"]
MORE_GOLDEN_NAMES = ["binom_small", "binom_groups", "binom_k0", "nb2_small", "nb2_ragged", "nb2_groups"]
CLASS_GOLDEN_NAMES = ["ordlog_small", "ordlog_ragged", "ordlog_binary", "ordlog_k0", "catlog_small", "catlog_ragged",
                      "catlog_one_class"]
