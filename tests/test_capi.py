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


def test_json_data_block_is_validated_like_a_stanc_model(tmp_path):
    """Host logic of the stanc-style constructor b200::glm_model(var_context&, glm_config): the data block is
    read through the reference's stan::json::json_data and checked with var_context::validate_dims, so the
    messages are the reference's.  Valid data then fails LOUDLY here (no GPU, no CPU fallback)."""
    import json
    from stan_b200 import stan_service
    from stan_b200.model import InvalidArgument
    if not stan_service.available():
        pytest.skip("libb200stan.so not built (needs the reference headers)")
    p = tmp_path / "d.json"
    p.write_text(json.dumps({"N": 3, "K": 2, "X": [[1, 2], [3, 4]], "y": [0, 1, 1]}))
    with pytest.raises(InvalidArgument, match="variable name=X.*dims declared=\\(3,2\\).*dims found=\\(2,2\\)"):
        stan_service.StanGLM.from_json(str(p), "bernoulli_logit")
    p.write_text(json.dumps({"N": 2, "K": 2, "X": [[1, 2], [3, 4]]}))
    with pytest.raises(InvalidArgument, match="variable name=y"):
        stan_service.StanGLM.from_json(str(p), "bernoulli_logit")
    p.write_text(json.dumps({"N": 2, "K": 2, "X": [[1, 2], [3, 4]], "y": [0, 1], "G": 2, "group": [1, 3]}))
    with pytest.raises(InvalidArgument, match="group"):
        stan_service.StanGLM.from_json(str(p), "bernoulli_logit")
    import torch
    if not torch.cuda.is_available():
        p.write_text(json.dumps({"N": 2, "K": 2, "X": [[1, 2], [3, 4]], "y": [0, 1]}))
        with pytest.raises(InvalidArgument, match="no CUDA device"):
            stan_service.StanGLM.from_json(str(p), "bernoulli_logit")


def test_stan_csv_reader_binding(tmp_path):
    """read_stan_csv is the reference's stan::io::stan_csv_reader::parse; a hand-written file in the layout the
    services write (header, adaptation block, draws, timing) must come back value for value."""
    from stan_b200 import stan_service
    if not stan_service.available():
        pytest.skip("libb200stan.so not built (needs the reference headers)")
    p = tmp_path / "x.csv"
    p.write_text("lp__,accept_stat__,alpha,beta.1,beta.2\n"
                 "# Adaptation terminated\n# Step size = 0.25\n# Diagonal elements of inverse mass matrix:\n"
                 "# 1.5, 2, 0.125\n"
                 "-1.5,0.9,0.1,0.2,0.3\n-2.5,1,0.4,0.5,0.6\n"
                 "# \n#  Elapsed Time: 0.1 seconds (Warm-up)\n#                0.2 seconds (Sampling)\n"
                 "#                0.3 seconds (Total)\n# \n")
    csv = stan_service.read_stan_csv(str(p))
    assert csv["header"] == ["lp__", "accept_stat__", "alpha", "beta[1]", "beta[2]"]
    assert np.array_equal(csv["samples"], [[-1.5, 0.9, 0.1, 0.2, 0.3], [-2.5, 1, 0.4, 0.5, 0.6]])
    assert csv["step_size"] == 0.25 and np.array_equal(csv["metric"], [1.5, 2, 0.125])


def test_desc_struct_layout_matches_the_header(tmp_path):
    """The ctypes mirrors of b200glm_desc (stan_b200/_capi.py) and glm_spec (oracle/oracle.py) must have the
    field offsets and size the C compiler gives the structs in include/b200glm.h and oracle/glm_oracle.h."""
    import ctypes as C
    import shutil
    import subprocess
    from oracle.oracle import GlmSpec
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    checks = (("b200glm_desc", os.path.join(ROOT, "include", "b200glm.h"), _capi.Desc),
              ("b200glm_nuts_config", os.path.join(ROOT, "include", "b200glm.h"), _capi.NutsConfig),
              ("b200glm_nuts_status", os.path.join(ROOT, "include", "b200glm.h"), _capi.NutsStatus),
              ("glm_spec", os.path.join(ROOT, "oracle", "glm_oracle.h"), GlmSpec))
    for struct, header, mirror in checks:
        fields = [f[0] for f in mirror._fields_]
        prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{header}"', "int main(void) {",
                f'  printf("%zu\\n", sizeof({struct}));']
        prog += [f'  printf("%zu\\n", offsetof({struct}, {f}));' for f in fields]
        prog += ["  return 0;", "}"]
        src, exe = tmp_path / f"{struct}.c", tmp_path / f"{struct}.bin"
        src.write_text("\n".join(prog))
        subprocess.run([cc, "-std=c11", "-o", str(exe), str(src)], check=True, capture_output=True)
        out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
        assert out[0] == C.sizeof(mirror), (struct, out[0], C.sizeof(mirror))
        for name, off in zip(fields, out[1:]):
            assert getattr(mirror, name).offset == off, (struct, name, off)
