"""CPU suite: the C-ABI library loads here (no GPU) and exports every symbol include/b200glm.h
declares; without a device the compute path fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import stan_b200
from stan_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200glm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200glm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    declared = header_symbols()
    assert len(declared) >= 15
    for s in declared:
        assert hasattr(L, s), s
    assert sorted(_capi.SYMBOLS) == declared
    assert b"sm_100a" in L.b200glm_version()


def test_no_cpu_fallback_without_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    d = stan_b200.make_glm_data("bernoulli_logit", 64, 3)
    with pytest.raises(stan_b200.CudaError):
        stan_b200.GLMModel("bernoulli_logit", d["X"], d["y"])


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (parity would be void)."""
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "stan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|glm_oracle", txt):
                    bad.append(os.path.join(root, f))
    assert not bad, bad


def test_argument_validation_host_side():
    with pytest.raises(stan_b200.InvalidArgument):
        stan_b200.GLMModel("bernoulli_logit", np.zeros((4, 2)), np.zeros(5, np.int32))
