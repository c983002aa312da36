"""Generates tests/golden/glm_function_golden.json from the REFERENCE itself (oracle/_ref, compiled from
/root/reference): the bare densities stan::math::{bernoulli_logit,poisson_log,normal_id}_glm_lp*f<propto>
and their adjoints, for every (propto, operands_are_var, sigma_is_var) combination.  Run in the build
container only:  python tests/golden/make_function_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle  # noqa: E402
from stan_b200 import make_glm_data  # noqa: E402

hx = lambda a: [float(v).hex() for v in np.atleast_1d(a)]
cases = []
for name, fam, N, K, G in [("f_bern", "bernoulli_logit", 257, 5, 0), ("f_bern_groups", "bernoulli_logit", 300, 3, 4),
                           ("f_pois", "poisson_log", 200, 4, 0), ("f_pois_groups", "poisson_log", 333, 2, 7),
                           ("f_norm", "normal_id", 150, 6, 0), ("f_norm_groups", "normal_id", 222, 3, 5)]:
    d = make_glm_data(fam, N, K, G, seed=99)
    rng = np.random.default_rng(5)
    a = 0.2 * rng.standard_normal(max(G, 1))
    b = 0.2 * rng.standard_normal(K)
    sigma = 1.3
    evals = []
    for propto in (0, 1):
        for ov in (0, 1):
            for sv in ((0, 1) if (fam == "normal_id" and ov) else (0,)):
                lp, da, db, ds = RefOracle.glm_function(fam, d["X"], d["y"], a, b, sigma, d["group"], G, propto, ov, sv)
                evals.append(dict(propto=propto, operands_are_var=ov, sigma_is_var=sv, lp=float(lp).hex(),
                                  d_alpha=hx(da), d_beta=hx(db), d_sigma=float(ds).hex()))
    cases.append(dict(name=name, family=fam, N=N, K=K, G=G, X=hx(d["X"].ravel(order="F")),
                      y=[float(v) if fam == "normal_id" else int(v) for v in d["y"]],
                      group=None if d["group"] is None else [int(v) for v in d["group"]],
                      alpha=hx(a), beta=hx(b), sigma=sigma, evals=evals))
out = os.path.join(ROOT, "tests", "golden", "glm_function_golden.json")
with open(out, "w") as f:
    json.dump(dict(source="stan-dev/stan@9048555 + math@2fdd3ed via oracle/_ref (ref_glm_function)", cases=cases), f)
print(out, os.path.getsize(out))
