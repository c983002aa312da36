"""Synthetic GLM data of the shapes BASELINE.json names.

One generator for both sides of every parity test: data is generated ONCE (host numpy
Philox for tests, device torch Philox for the HBM-scale bench) and the very same buffers
are handed to the CUDA path and to the CPU checker, because host libm and CUDA
transcendental last bits differ and independently regenerated X would not be identical.
"""
import numpy as np

SEED = 20261017


def _rng(seed, stream=0):
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def make_glm_data(family, N, K, G=0, seed=SEED, alpha_true=0.3, n_classes=4):
    """Returns dict(X (N,K) F-order float64, y, group (1-based int32 or None), truth); binomial_logit adds
    `trials`, the class models (ordered_logistic, categorical_logit: oracle-side only so far) `n_classes`."""
    r = _rng(seed, 0)
    X = np.asfortranarray(r.standard_normal((K, N)).T) if N * K else np.zeros((N, K), order="F")
    beta = _rng(seed, 1).standard_normal(K) / np.sqrt(max(K, 1))
    eta = X @ beta if K else np.zeros(N)
    group = None
    a_true = None
    if G:
        rg = _rng(seed, 2)
        group = rg.integers(1, G + 1, size=N, dtype=np.int32)
        a_true = 0.5 * rg.standard_normal(G)
        eta = eta + a_true[group - 1]
    else:
        eta = eta + alpha_true
    ry = _rng(seed, 3)
    if family == "bernoulli_logit":
        y = (ry.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.int32)
    elif family == "poisson_log":
        y = ry.poisson(np.exp(np.clip(0.5 * eta + 0.5, -20, 5))).astype(np.int32)
    elif family == "normal_id":
        y = eta + ry.standard_normal(N)
    elif family == "binomial_logit":
        trials = ry.integers(0, 41, size=N, dtype=np.int32)          # includes empty populations
        y = ry.binomial(trials, 1.0 / (1.0 + np.exp(-eta))).astype(np.int32)
    elif family == "neg_binomial_2_log":
        mu, phi_true = np.exp(np.clip(0.5 * eta + 0.5, -20, 5)), 2.0
        y = ry.poisson(ry.gamma(phi_true, mu / phi_true)).astype(np.int32)   # gamma-poisson mixture
    elif family == "ordered_logistic":
        cuts = np.linspace(-1.5, 1.5, max(n_classes - 1, 0))
        lat = eta - alpha_true + ry.logistic(size=N)
        y = (1 + (lat[:, None] > cuts[None, :]).sum(axis=1)).astype(np.int32)
    elif family == "categorical_logit":
        Bc = _rng(seed, 4).standard_normal((K, n_classes)) / np.sqrt(max(K, 1))
        lin = (X @ Bc if K else np.zeros((N, n_classes))) + 0.2 * np.arange(n_classes)
        gum = ry.gumbel(size=(N, n_classes))
        y = (1 + np.argmax(lin + gum, axis=1)).astype(np.int32) if N else np.zeros(0, np.int32)
    else:
        raise ValueError(family)
    out = dict(family=family, X=X, y=y, group=group, G=G,
               truth=dict(alpha=alpha_true, beta=beta, a=a_true))
    if family == "binomial_logit":
        out["trials"] = trials
    if family in ("ordered_logistic", "categorical_logit"):
        out["n_classes"] = n_classes
    return out


def theta_points(P, seed=SEED, n_random=1, scale=0.1):
    """theta = 0 and random N(0, scale^2) points (evaluate at >= 3 thetas;
    the near-mode point comes from a short warm-up in the tests that need it)."""
    pts = [np.zeros(P)]
    r = _rng(seed, 7)
    for _ in range(n_random):
        pts.append(scale * r.standard_normal(P))
    return pts


# ---------------------------------------------------------------------------------------------
# HBM-scale synthetic data, generated on the device that will hold it (bench.py, multi-GPU tests)
# ---------------------------------------------------------------------------------------------
def shard_rows(N_total, rank, world):
    """Contiguous row block [r0, r1) of rank `rank` of `world` (SURVEY 8e partitioning)."""
    per = (N_total + world - 1) // world
    r0 = min(rank * per, N_total)
    return r0, min(r0 + per, N_total)


def shard_rows_weighted(N_total, weights, rank, align=32):
    """Row block of `rank` when the rows are split in proportion to `weights` (one positive number per rank, e.g.
    a measured per-GPU memory bandwidth: a row-sharded step waits for its slowest shard, so a slower GPU should
    hold fewer rows).  Boundaries are multiples of `align` rows (the panel height); the blocks are contiguous,
    disjoint and cover [0, N_total)."""
    w = np.asarray(weights, dtype=np.float64)
    if w.ndim != 1 or w.size < 1 or not np.all(w > 0) or not np.all(np.isfinite(w)):
        raise ValueError("weights must be positive and finite, one per rank")
    cum = np.concatenate([[0.0], np.cumsum(w)]) / w.sum()
    bounds = [min(N_total, int(round(N_total * c / align)) * align) for c in cum]
    bounds[0], bounds[-1] = 0, N_total
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds[rank], bounds[rank + 1]


def make_logistic_shard(torch, dev, N_total, K, rank=0, world=1, seed=SEED, block=1_000_000, alpha_true=0.3,
                        rows=None):
    """Rows shard_rows(N_total, rank, world) of the synthetic logistic-regression problem,
    column-major fp64 X (stored as a (K, n) torch tensor) and int32 y, generated with torch's
    Philox on `dev`.  Every `block`-row block has its own seed, so any sharding of the same
    (N_total, K, seed) reproduces the same global matrix on the same device type.  rows=(r0, r1) overrides the
    equal split (shard_rows_weighted)."""
    r0, r1 = rows if rows is not None else shard_rows(N_total, rank, world)
    n = r1 - r0
    g = torch.Generator(device=dev)
    beta = torch.from_numpy(_rng(seed, 1).standard_normal(K) / np.sqrt(max(K, 1))).to(dev)
    X = torch.empty((K, n), device=dev, dtype=torch.float64)
    y = torch.empty(n, device=dev, dtype=torch.int32)
    for b0 in range((r0 // block) * block, r1, block):
        g.manual_seed(seed * 1000 + b0 // block)
        xb = torch.randn((K, block), generator=g, device=dev, dtype=torch.float64)
        ub = torch.rand(block, generator=g, device=dev, dtype=torch.float64)
        lo, hi = max(b0, r0), min(b0 + block, r1)
        xs = xb[:, lo - b0:hi - b0]
        X[:, lo - r0:hi - r0] = xs
        eta = alpha_true + beta @ xs
        y[lo - r0:hi - r0] = (ub[lo - b0:hi - b0] < torch.sigmoid(eta)).to(torch.int32)
        del xb, ub
    return X, y, r0, r1


def make_shard(torch, dev, family, N_total, K, G=0, rank=0, world=1, seed=SEED, block=1_000_000, alpha_true=0.3):
    """Rows shard_rows(N_total, rank, world) of the synthetic problem for any family / grouping
    (SURVEY 8d generator), on `dev`: returns X (K, n) fp64 = column-major N x K, y (int32 or fp64),
    group (int32, 1-based, or None), r0, r1.  Per-block seeds as in make_logistic_shard.
    (binomial_logit also needs the population sizes: use make_shard_ex.)"""
    X, y, group, trials, r0, r1 = make_shard_ex(torch, dev, family, N_total, K, G, rank, world, seed, block, alpha_true)
    return X, y, group, r0, r1


def make_shard_ex(torch, dev, family, N_total, K, G=0, rank=0, world=1, seed=SEED, block=1_000_000, alpha_true=0.3,
                  rows=None):
    """make_shard plus the binomial population sizes: returns X, y, group, trials (int32 or None), r0, r1.
    rows=(r0, r1) overrides the equal split."""
    if family == "bernoulli_logit" and G == 0:
        X, y, r0, r1 = make_logistic_shard(torch, dev, N_total, K, rank, world, seed, block, alpha_true, rows)
        return X, y, None, None, r0, r1
    r0, r1 = rows if rows is not None else shard_rows(N_total, rank, world)
    n = r1 - r0
    g = torch.Generator(device=dev)
    beta = torch.from_numpy(_rng(seed, 1).standard_normal(K) / np.sqrt(max(K, 1))).to(dev)
    a_true = torch.from_numpy(0.5 * _rng(seed, 2).standard_normal(max(G, 1))).to(dev)
    X = torch.empty((K, n), device=dev, dtype=torch.float64)
    y = torch.empty(n, device=dev, dtype=torch.float64 if family == "normal_id" else torch.int32)
    group = torch.empty(n, device=dev, dtype=torch.int32) if G else None
    trials = torch.empty(n, device=dev, dtype=torch.int32) if family == "binomial_logit" else None
    for b0 in range((r0 // block) * block, r1, block):
        g.manual_seed(seed * 1000 + b0 // block)
        xb = torch.randn((K, block), generator=g, device=dev, dtype=torch.float64)
        ub = torch.rand(block, generator=g, device=dev, dtype=torch.float64)
        gb = torch.randint(1, max(G, 1) + 1, (block,), generator=g, device=dev, dtype=torch.int32)
        lo, hi = max(b0, r0), min(b0 + block, r1)
        X[:, lo - r0:hi - r0] = xb[:, lo - b0:hi - b0]
        # y is drawn for the WHOLE block and then sliced, so every sharding sees the same random stream
        eta = beta @ xb
        eta = eta + (a_true[(gb - 1).long()] if G else alpha_true)
        if family == "bernoulli_logit":
            yb = (ub < torch.sigmoid(eta)).to(torch.int32)
        elif family == "poisson_log":
            yb = torch.poisson(torch.exp(torch.clamp(0.5 * eta + 0.5, -20, 5)), generator=g).to(torch.int32)
        elif family == "binomial_logit":
            tb = torch.randint(0, 41, (block,), generator=g, device=dev, dtype=torch.int32)
            # successes: sum of `trials` Bernoulli draws, 40 uniform columns at a time
            u = torch.rand((40, block), generator=g, device=dev, dtype=torch.float32)
            hit = (u < torch.sigmoid(eta).to(torch.float32)) & (torch.arange(40, device=dev)[:, None] < tb[None, :])
            yb = hit.sum(0).to(torch.int32)
            trials[lo - r0:hi - r0] = tb[lo - b0:hi - b0]
            del u, hit, tb
        elif family == "neg_binomial_2_log":
            mu = torch.exp(torch.clamp(0.5 * eta + 0.5, -20, 5))
            # gamma(shape 2, mean mu) as a sum of two exponentials (keeps every draw on generator g): phi = 2 mixture
            e2 = torch.rand((2, block), generator=g, device=dev, dtype=torch.float64).clamp_min(1e-300)
            lam = -(mu / 2.0) * torch.log(e2).sum(0)
            yb = torch.poisson(lam, generator=g).to(torch.int32)
        else:
            yb = eta + torch.randn(block, generator=g, device=dev, dtype=torch.float64)
        y[lo - r0:hi - r0] = yb[lo - b0:hi - b0]
        if G:
            group[lo - r0:hi - r0] = gb[lo - b0:hi - b0]
        del xb, ub, gb
    return X, y, group, trials, r0, r1
