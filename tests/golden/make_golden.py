"""Generate tests/golden/glm_ref_golden.json from the COMPILED REFERENCE (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py          # glm_ref_golden.json (bernoulli_logit, poisson_log, normal_id)
    python tests/golden/make_golden.py more     # glm_more_families_golden.json (binomial_logit, neg_binomial_2_log)
    python tests/golden/make_golden.py classes  # glm_class_models_golden.json (ordered_logistic, categorical_logit)
Every case stores its full inputs (X, y, group, theta as float.hex strings, so they are exact)
and the reference's outputs: stan::model::log_prob_grad<propto,jacobian> value + gradient,
Model::log_prob<propto,jacobian>(double) values, and one expl_leapfrog step.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle, num_params  # noqa: E402
from stan_b200.synth import make_glm_data  # noqa: E402

CASES = [
    # name, family, N, K, G
    ("bern_small", "bernoulli_logit", 64, 5, 0),
    ("bern_ragged", "bernoulli_logit", 153, 7, 0),     # N not a multiple of the 32-row panel
    ("bern_wide", "bernoulli_logit", 40, 33, 0),
    ("bern_groups", "bernoulli_logit", 97, 4, 6),
    ("pois_small", "poisson_log", 64, 5, 0),
    ("pois_groups", "poisson_log", 130, 3, 11),
    ("norm_small", "normal_id", 64, 5, 0),
    ("norm_ragged", "normal_id", 71, 9, 0),
    ("norm_groups", "normal_id", 90, 2, 5),
    ("bern_k1", "bernoulli_logit", 33, 1, 0),
    ("bern_k0", "bernoulli_logit", 20, 0, 0),          # zero attributes
    ("pois_n1", "poisson_log", 1, 3, 0),
]


MORE_CASES = [
    ("binom_small", "binomial_logit", 64, 5, 0),
    ("binom_groups", "binomial_logit", 97, 4, 6),       # ragged N, a[group]
    ("binom_k0", "binomial_logit", 20, 0, 0),           # empty weight vector: the reference returns 0 (size_zero)
    ("nb2_small", "neg_binomial_2_log", 64, 5, 0),
    ("nb2_ragged", "neg_binomial_2_log", 71, 9, 0),
    ("nb2_groups", "neg_binomial_2_log", 130, 3, 11),
]


CLASS_CASES = [
    # name, family, N, K, n_classes  (oracle-side goldens: these two GLMs are not built on the device yet)
    ("ordlog_small", "ordered_logistic", 64, 5, 4),
    ("ordlog_ragged", "ordered_logistic", 153, 7, 6),
    ("ordlog_binary", "ordered_logistic", 40, 3, 2),
    ("ordlog_k0", "ordered_logistic", 30, 0, 3),
    ("catlog_small", "categorical_logit", 64, 5, 4),
    ("catlog_ragged", "categorical_logit", 97, 9, 3),
    ("catlog_one_class", "categorical_logit", 20, 2, 1),       # N_classes == 1: the reference returns 0
]


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def main(cases=CASES, filename="glm_ref_golden.json", seed0=4242):
    out = {"reference": RefOracle.lib().ref_oracle_version().decode(), "cases": []}
    for name, fam, N, K, G in cases:
        n_classes = 0
        if fam in ("ordered_logistic", "categorical_logit"):
            n_classes, G = G, 0                      # CLASS_CASES carry the class count in the fifth field
        d = make_glm_data(fam, N, K, G, seed=seed0 + len(out["cases"]), **({"n_classes": n_classes} if n_classes else {}))
        kw = {"trials": d["trials"]} if "trials" in d else {}
        if n_classes:
            kw["n_classes"] = n_classes
        ro = RefOracle(fam, d["X"], d["y"], d["group"], G, **kw)
        P = num_params(fam, K, G, n_classes)
        assert P == ro.P
        rng = np.random.default_rng(99 + len(out["cases"]))
        thetas = [np.zeros(P), 0.3 * rng.standard_normal(P), 1.5 * rng.standard_normal(P)]
        if fam == "bernoulli_logit" and K > 0:
            big = np.zeros(P)
            big[-K:] = 12.0   # drives |ytheta| beyond the cutoff of 20 on many rows (both branches)
            thetas.append(big)
        evals = []
        for th in thetas:
            e = {"theta": hx(th), "lp_grad": {}, "lp_double": {}}
            for propto in (1, 0):
                for jac in (1, 0):
                    lp, g = ro.log_prob_grad(th, propto, jac)
                    e["lp_grad"][f"{propto}{jac}"] = {"lp": float(lp).hex(), "grad": hx(g)}
                    e["lp_double"][f"{propto}{jac}"] = float(ro.log_prob(th, propto, jac)).hex()
            evals.append(e)
        # one leapfrog step of the reference integrator from theta[1]
        q0 = thetas[1]
        p0 = rng.standard_normal(P)
        im = np.exp(0.3 * rng.standard_normal(P))
        lp0, g0 = ro.log_prob_grad(q0, 1, 1)
        q1, p1, g1, V1 = ro.leapfrog(0.01, im, q0, p0, -g0, -lp0)
        case = dict(name=name, family=fam, N=N, K=K, G=G,
                    X=hx(np.asarray(d["X"]).ravel(order="F")), y=[float(v) for v in d["y"]],
                    group=None if d["group"] is None else [int(v) for v in d["group"]],
                    trials=[int(v) for v in d["trials"]] if "trials" in d else None, n_classes=n_classes,
                    evals=evals,
                    leapfrog=dict(eps=0.01, inv_metric=hx(im), q0=hx(q0), p0=hx(p0), g0=hx(-g0), V0=float(-lp0).hex(),
                                  q1=hx(q1), p1=hx(p1), g1=hx(g1), V1=float(V1).hex()))
        out["cases"].append(case)
        print(name, "P=", P, "lp(theta1)=", ro.log_prob_grad(thetas[1])[0])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), filename)
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "more":
        main(MORE_CASES, "glm_more_families_golden.json", seed0=5151)
    elif len(sys.argv) > 1 and sys.argv[1] == "classes":
        main(CLASS_CASES, "glm_class_models_golden.json", seed0=6161)
    else:
        main()
