"""Small driver for ncu captures: builds one model on cuda:0 and runs a few device-resident leapfrog steps, so that a
capture filtered on `glm_` sees only this repo's kernels in steady state.
    ncu --set full --clock-control none -k regex:glm_ --launch-skip 4 --launch-count 2 -o out python tools/ncu_target.py cfg2
targets: multi4 (4 chains, 4M x 100) | cfg2 | cfg2shard (1.25M rows) | cfg4shard (6.25M rows, 1000 groups) | wide (1M x 1000) | cfg3 (1024 chains) |
         ordlog | catlog"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import GLMModel  # noqa: E402
from stan_b200.synth import make_shard_ex  # noqa: E402

T = {"cfg2": ("bernoulli_logit", 10_000_000, 100, 0), "cfg2shard": ("bernoulli_logit", 1_250_000, 100, 0),
     "cfg4shard": ("poisson_log", 6_250_000, 50, 1000), "wide": ("bernoulli_logit", 1_000_000, 1000, 0),
     "cfg3": ("normal_id", 1_000_000, 200, 0), "ordlog": ("ordered_logistic", 10_000_000, 100, 0),
     "catlog": ("categorical_logit", 10_000_000, 100, 0), "multi4": ("bernoulli_logit", 4_000_000, 100, 0)}
name = sys.argv[1]
fam, N, K, G = T[name]
dev = torch.device("cuda", 0)
ncls = 0
if fam in ("ordered_logistic", "categorical_logit"):
    ncls = 5 if fam == "ordered_logistic" else 4
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn((K, N), generator=g, device=dev, dtype=torch.float64)
    y = torch.randint(1, ncls + 1, (N,), generator=g, device=dev, dtype=torch.int32)
    grp = tr = None
else:
    X, y, grp, tr, _, _ = make_shard_ex(torch, dev, fam, N, K, G, 0, 1)
m = GLMModel(fam, X.data_ptr(), y.data_ptr(), grp.data_ptr() if G else None, G, data_on_device=True, N=N, K=K, ldx=N,
             n_classes=ncls)
del X, y
rng = np.random.default_rng(11)
if name in ("cfg3", "multi4"):
    C = 1024 if name == "cfg3" else 4     # multi4: four chains per pass on the FMA path (glm_multi_kernel)
    m.batch_reserve(C)
    q0, p0 = 0.05 * rng.standard_normal((C, m.P)), rng.standard_normal((C, m.P))
    lp0, g0, st = m.log_prob_grad_batched(q0)
    m.set_state_batched(q0, p0, -g0, -lp0)
    for _ in range(4):
        m.leapfrog_batched_async(C, 1e-5)
    m.batch_sync()
else:
    q0, p0 = 0.05 * rng.standard_normal(m.P), rng.standard_normal(m.P)
    lp0, g0 = m.log_prob_grad(q0)
    m.set_state(q0, p0, -g0, -lp0)
    for _ in range(8):
        m.leapfrog_async(1e-4)
    m.sync()
m.close()
print("done", name)
