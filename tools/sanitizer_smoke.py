"""One small call of every kernel family, for `compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py`.

No torch, no oracle: ctypes + numpy only, shapes chosen so that every kernel sees a ragged last panel and (where it has
one) its padded-column path: narrow kernel (scalar intercept, groups fused and unfused, families 0-4), wide kernel, class
kernels, batched DMMA (normal and row-split), few-chain FMA kernel (one and two passes), batched leapfrog, device-side NUTS
rounds.  Prints 'SANITIZER-SMOKE-DONE' at the end; the sanitizer's own report says whether any access was out of bounds."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import GLMModel, make_glm_data  # noqa: E402


def single(fam, N, K, G=0, **kw):
    d = make_glm_data(fam, N, K, G, **({"n_classes": kw["n_classes"]} if "n_classes" in kw else {}))
    extra = dict(kw)
    if fam == "binomial_logit":
        extra["trials"] = d["trials"]
    m = GLMModel(fam, d["X"], d["y"], d["group"], G, **extra)
    th = 0.1 * np.random.default_rng(1).standard_normal(m.P)
    lp, g = m.log_prob_grad(th)
    m.set_state(th, np.ones(m.P), -g, -lp)
    m.leapfrog(1e-3)
    assert np.isfinite(lp)
    m.close()
    print("single", fam, N, K, G, "ok", flush=True)


def batched(fam, N, K, lanes):
    d = make_glm_data(fam, N, K)
    extra = {"trials": d["trials"]} if fam == "binomial_logit" else {}
    m = GLMModel(fam, d["X"], d["y"], **extra)
    m.batch_reserve(max(lanes))
    rng = np.random.default_rng(2)
    for n in lanes:
        th = 0.1 * rng.standard_normal((n, m.P))
        lp, g, st = m.log_prob_grad_batched(th)
        assert not st.any() and np.all(np.isfinite(lp))
        m.set_state_batched(th, rng.standard_normal((n, m.P)), -g, -lp)
        m.leapfrog_batched(np.full(n, 1e-3))
    m.close()
    print("batched", fam, N, K, lanes, "ok", flush=True)


def device_nuts(fam, N, K, chains):
    from stan_b200 import stan_service
    if not stan_service.available():
        return
    d = make_glm_data(fam, N, K)
    m = stan_service.StanGLM(fam, d["X"], d["y"])
    r = m.nuts_device(num_chains=chains, seed=3, num_warmup=30, num_samples=10, delta=0.8, stepsize_jitter=0.2)
    assert np.all(np.isfinite(r["draws"]))
    m.close()
    print("device nuts", fam, N, K, chains, "ok", flush=True)


if __name__ == "__main__":
    batched("bernoulli_logit", 3_003, 100, (70, 16, 4, 7))      # two chain blocks, row-split, few-chain one / two passes
    batched("binomial_logit", 1_030, 17, (33, 9, 3))
    batched("normal_id", 2_001, 200, (40, 12))                  # DMMA normal mode, row-split (K > 128: no few-chain kernel)
    batched("poisson_log", 31, 3, (2,))                         # a single partial panel
    device_nuts("bernoulli_logit", 2_000, 6, 5)
    single("bernoulli_logit", 3_001, 20)
    single("poisson_log", 2_050, 7, 13)                         # fused group path
    single("normal_id", 1_999, 33)
    single("binomial_logit", 2_047, 18)
    single("neg_binomial_2_log", 1_025, 9)
    single("bernoulli_logit", 1_500, 300)                       # wide kernel
    single("ordered_logistic", 2_001, 12, n_classes=5)
    single("categorical_logit", 1_777, 9, n_classes=4)
    print("SANITIZER-SMOKE-DONE", flush=True)
