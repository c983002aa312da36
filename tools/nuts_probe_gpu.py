"""Diagnostic: per-chain work distribution of the batched NUTS driver (not part of the test suite)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from stan_b200 import make_glm_data, stan_service
N, K, C, W, S = (int(a) for a in sys.argv[1:6])
d = make_glm_data("normal_id", N, K)
m = stan_service.StanGLM("normal_id", d["X"], d["y"], n_slots=16)
t0 = time.time()
res = m.nuts_batched(num_chains=C, seed=4711, num_warmup=W, num_samples=S, delta=0.8)
print("N K C", N, K, C, "wall", round(time.time() - t0, 1), "batches", res["batches"], "lanes", res["lanes"])
per_chain = res["warm_leapfrogs"] + res["draws"][:, :, 4].sum(axis=1)
print("per-chain leapfrogs pct 50/90/99/max", [float(np.percentile(per_chain, q)) for q in (50, 90, 99, 100)])
print("stepsize pct 0/1/50/100", [float(np.percentile(res["stepsize"], q)) for q in (0, 1, 50, 100)])
print("chains mean treedepth>=8:", int((res["draws"][:, :, 3].mean(axis=1) >= 8).sum()), "divergent", int(res["draws"][:, :, 5].sum()))
wd = res["warmup_draws"]
nb = max(W // 10, 1)
print("warmup n_leapfrog mean by block:", [round(float(wd[:, i:i + nb, 4].mean()), 1) for i in range(0, W, nb)])
print("warmup n_leapfrog max by block:", [float(wd[:, i:i + nb, 4].max()) for i in range(0, W, nb)])
sig = res["draws"][:, :, -1]
print("sigma post mean range over chains", float(sig.mean(axis=1).min()), float(sig.mean(axis=1).max()))
bad = np.argsort(per_chain)[-3:]
for c in bad:
    print("chain", int(c), "leapfrogs", float(per_chain[c]), "stepsize", float(res["stepsize"][c]), "warmup lp first/last",
          float(wd[c, 0, 0]), float(wd[c, -1, 0]), "n_leapfrog warmup head", wd[c, :12, 4].tolist())
m.close()
