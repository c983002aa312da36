"""GPU parity suite: the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): log_prob and gradient within 1e-10 relative of the reference's
stan::model::log_prob_grad.  Gradient entries are scaled by max(|g_k|, ||g||_inf) (conftest.rel_err_vec).
"""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, rel_err, rel_err_vec, unhex
import stan_b200
from stan_b200 import GLMModel, make_glm_data, theta_points

pytestmark = pytest.mark.gpu
TOL = 1e-10


def oracle_for(fam, d, G=0, **pri):
    from oracle.oracle import PortOracle, RefOracle
    cls = RefOracle if RefOracle.available() else PortOracle
    return cls(fam, d["X"], d["y"], d["group"], G, **pri)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_cuda_matches_reference_golden(golden, name):
    c = golden[name]
    m = GLMModel(c["family"], c["X"], c["y"], c["group"], c["G"])
    assert m.num_params_r() == len(c["evals"][0]["theta"])
    for e in c["evals"]:
        th = unhex(e["theta"])
        for key, ref in e["lp_grad"].items():
            propto, jac = int(key[0]), int(key[1])
            lp, g = m.log_prob_grad(th, propto, jac)
            assert rel_err(lp, float.fromhex(ref["lp"])) < TOL, (name, key, lp, float.fromhex(ref["lp"]))
            assert rel_err_vec(g, unhex(ref["grad"])) < TOL, (name, key)
        for key, ref in e["lp_double"].items():
            propto, jac = int(key[0]), int(key[1])
            assert rel_err(m.log_prob(th, propto, jac), float.fromhex(ref)) < TOL, (name, key)
    m.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_cuda_leapfrog_matches_reference_golden(golden, name):
    c = golden[name]
    lf = c["leapfrog"]
    m = GLMModel(c["family"], c["X"], c["y"], c["group"], c["G"])
    m.set_state(unhex(lf["q0"]), unhex(lf["p0"]), unhex(lf["g0"]), float.fromhex(lf["V0"]))
    q, p, g, V = m.leapfrog(lf["eps"], unhex(lf["inv_metric"]))
    assert rel_err_vec(q, unhex(lf["q1"])) < TOL
    assert rel_err_vec(p, unhex(lf["p1"])) < TOL
    assert rel_err_vec(g, unhex(lf["g1"])) < TOL
    assert rel_err(V, float.fromhex(lf["V1"])) < TOL
    m.close()


SHAPES = [
    # family, N, K, G  -- config 1 shape, ragged N, K around the accumulator-slot boundaries, groups
    ("bernoulli_logit", 10_000, 20, 0),
    ("bernoulli_logit", 50_001, 100, 0),
    ("bernoulli_logit", 4_097, 13, 0),
    ("bernoulli_logit", 3_000, 57, 0),
    ("bernoulli_logit", 2_000, 200, 0),
    ("bernoulli_logit", 1_000, 256, 0),
    ("poisson_log", 60_000, 50, 0),
    ("poisson_log", 80_000, 50, 1000),
    ("poisson_log", 5_000, 8, 3000),      # G > SMEM_A_MAX_GROUPS: a[] read from global
    ("normal_id", 40_000, 200, 0),
    ("normal_id", 9_999, 31, 17),
    ("bernoulli_logit", 7_777, 5, 12),
]


@pytest.mark.parametrize("fam,N,K,G", SHAPES)
def test_cuda_matches_oracle_live(fam, N, K, G):
    d = make_glm_data(fam, N, K, G)
    orc = oracle_for(fam, d, G)
    m = GLMModel(fam, d["X"], d["y"], d["group"], G)
    assert m.num_params_r() == orc.P
    for th in theta_points(m.P, n_random=2, scale=0.1):
        lp, g = m.log_prob_grad(th)
        lp_r, g_r = orc.log_prob_grad(th)
        assert rel_err(lp, lp_r) < TOL, (lp, lp_r)
        assert rel_err_vec(g, g_r) < TOL
        assert rel_err(m.log_prob(th, False, True), orc.log_prob(th, False, True)) < TOL
    m.close()


WIDE_SHAPES = [
    # family, N, K, G, flags -- the wide-matrix kernel (16-row panels split over the CTA): K > 256 selects it,
    # flags=1 (B200GLM_FLAG_FORCE_WIDE) runs it on narrow shapes too (sub-panel/warp-ownership edge cases)
    ("bernoulli_logit", 3_000, 300, 0, 0),       # KC=64, J=5
    ("bernoulli_logit", 20_011, 1000, 0, 0),     # config 5's K: KC=128, J=8
    ("normal_id", 5_000, 511, 0, 0),             # Cpad = 512: last KC=64 shape, y in the last column
    ("normal_id", 5_000, 512, 0, 0),             # first KC=128 shape
    ("poisson_log", 8_000, 520, 100, 0),         # groups; y and group id in the last sub-panel
    ("poisson_log", 4_000, 639, 50, 0),          # K+1 = 640: y and group id land in DIFFERENT sub-panels
    ("bernoulli_logit", 2_500, 1400, 0, 0),      # two sub-panels per warp
    ("normal_id", 1_201, 2_500, 0, 0),           # three sub-panels per warp
    ("bernoulli_logit", 10_000, 20, 0, 1),
    ("bernoulli_logit", 4_097, 13, 7, 1),
    ("poisson_log", 5_000, 8, 3000, 1),          # a[] read from global memory
    ("normal_id", 9_999, 100, 17, 1),
    ("normal_id", 15, 3, 0, 1),                  # a single partial panel
    ("bernoulli_logit", 16 * 148 * 3 + 5, 70, 0, 1),
]


@pytest.mark.parametrize("fam,N,K,G,flags", WIDE_SHAPES)
def test_wide_kernel_matches_oracle_live(fam, N, K, G, flags):
    d = make_glm_data(fam, N, K, G)
    orc = oracle_for(fam, d, G)
    m = GLMModel(fam, d["X"], d["y"], d["group"], G, flags=flags)
    assert m.num_params_r() == orc.P
    sc = 0.1 if K <= 300 else 0.03
    for th in theta_points(m.P, n_random=2, scale=sc):
        lp, g = m.log_prob_grad(th)
        lp_r, g_r = orc.log_prob_grad(th)
        assert rel_err(lp, lp_r) < TOL, (lp, lp_r)
        assert rel_err_vec(g, g_r) < TOL
        assert rel_err(m.log_prob(th, False, True), orc.log_prob(th, False, True)) < TOL
    a = m.log_prob_grad(th)
    b = m.log_prob_grad(th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])      # deterministic
    m.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_wide_kernel_matches_reference_golden(golden, name):
    c = golden[name]
    m = GLMModel(c["family"], c["X"], c["y"], c["group"], c["G"], flags=1)
    for e in c["evals"]:
        th = unhex(e["theta"])
        for key, ref in e["lp_grad"].items():
            lp, g = m.log_prob_grad(th, int(key[0]), int(key[1]))
            assert rel_err(lp, float.fromhex(ref["lp"])) < TOL, (name, key)
            assert rel_err_vec(g, unhex(ref["grad"])) < TOL, (name, key)
    lf = c["leapfrog"]
    m.set_state(unhex(lf["q0"]), unhex(lf["p0"]), unhex(lf["g0"]), float.fromhex(lf["V0"]))
    q, p, g, V = m.leapfrog(lf["eps"], unhex(lf["inv_metric"]))
    assert rel_err_vec(q, unhex(lf["q1"])) < TOL and rel_err_vec(p, unhex(lf["p1"])) < TOL
    assert rel_err_vec(g, unhex(lf["g1"])) < TOL and rel_err(V, float.fromhex(lf["V1"])) < TOL
    m.close()


def test_wide_kernel_hbm_scale_additivity():
    """K = 1000 (config 5's width) at a size well beyond L2: the likelihood part of 8 row shards sums
    to the whole (size-independent 'checksum of checksums'), and one shard is checked against the oracle."""
    import torch
    from oracle.oracle import PortOracle
    from stan_b200.synth import make_logistic_shard
    N, K, S = 800_000, 1000, 8
    dev = torch.device("cuda", 0)
    X, y, _, _ = make_logistic_shard(torch, dev, N, K, block=100_000)
    m = GLMModel("bernoulli_logit", X.data_ptr(), y.data_ptr(), data_on_device=True, N=N, K=K, ldx=N)
    th = 0.02 * np.random.default_rng(11).standard_normal(K + 1)
    lp, g = m.log_prob_grad(th)
    prior = PortOracle("bernoulli_logit", np.zeros((0, K)), np.zeros(0, np.int32))
    lp0, g0 = prior.log_prob_grad(th)
    lp_sum, g_sum = lp0, g0.copy()
    n = N // S
    for s in range(S):
        ms = GLMModel("bernoulli_logit", X.data_ptr() + 8 * s * n, y.data_ptr() + 4 * s * n, data_on_device=True,
                      N=n, K=K, ldx=N)
        lps, gs = ms.log_prob_grad(th)
        if s == 3:
            po = PortOracle("bernoulli_logit", X[:, s * n:(s + 1) * n].cpu().numpy().T, y[s * n:(s + 1) * n].cpu().numpy())
            lp_r, g_r = po.log_prob_grad(th)
            assert rel_err(lps, lp_r) < TOL and rel_err_vec(gs, g_r) < TOL
        lp_sum += lps - lp0
        g_sum += gs - g0
        ms.close()
    assert rel_err(lp, lp_sum) < 1e-12 and rel_err_vec(g, g_sum) < 1e-12
    m.close()


def test_deterministic_bitwise():
    d = make_glm_data("bernoulli_logit", 200_000, 100)
    m = GLMModel("bernoulli_logit", d["X"], d["y"])
    th = theta_points(m.P)[1]
    a = [m.log_prob_grad(th) for _ in range(3)]
    for lp, g in a[1:]:
        assert lp == a[0][0] and np.array_equal(g, a[0][1])
    m.close()


def test_near_mode_gradient_cancellation():
    """At the posterior mode the gradient entries are sums of O(N) terms cancelling to ~0."""
    d = make_glm_data("bernoulli_logit", 100_000, 20)
    orc = oracle_for("bernoulli_logit", d)
    m = GLMModel("bernoulli_logit", d["X"], d["y"])
    th = np.zeros(m.P)
    for _ in range(25):   # Newton-free: a few damped gradient steps get close enough to the mode
        lp, g = orc.log_prob_grad(th)
        th = th + 4.0 * g / d["X"].shape[0]
    lp, g = m.log_prob_grad(th)
    lp_r, g_r = orc.log_prob_grad(th)
    assert rel_err(lp, lp_r) < TOL
    # against the sum of |terms| (~N/4) the absolute error must be tiny even though g ~ 0
    assert np.max(np.abs(g - g_r)) < 1e-10 * d["X"].shape[0]
    m.close()


def test_leapfrog_trajectory_matches_oracle():
    """20 consecutive device-resident leapfrog steps == 20 oracle steps (state never re-uploaded)."""
    from oracle.oracle import PortOracle
    d = make_glm_data("poisson_log", 20_000, 10)
    po = PortOracle("poisson_log", d["X"], d["y"])
    m = GLMModel("poisson_log", d["X"], d["y"])
    rng = np.random.default_rng(5)
    q = 0.05 * rng.standard_normal(m.P)
    p = rng.standard_normal(m.P)
    im = np.exp(0.2 * rng.standard_normal(m.P))
    lp, g = po.log_prob_grad(q)
    g, V = -g, -lp
    m.set_state(q, p, g, V)
    qd, pd, gd, Vd = m.leapfrog(1e-3, im)
    q, p, g, V = po.leapfrog(1e-3, im, q, p, g, V)
    for _ in range(19):
        qd, pd, gd, Vd = m.leapfrog(1e-3)           # inv_metric NULL = unchanged
        q, p, g, V = po.leapfrog(1e-3, im, q, p, g, V)
    assert rel_err_vec(qd, q) < 1e-9 and rel_err_vec(pd, p) < 1e-9 and rel_err(Vd, V) < 1e-9
    # negative eps (backward tree, base_nuts.hpp:254-255) returns along the trajectory
    m.set_state(q, -p * 0 + p, g, V)
    for _ in range(20):
        qd, pd, gd, Vd = m.leapfrog(-1e-3)
        q, p, g, V = po.leapfrog(-1e-3, im, q, p, g, V)
    assert rel_err_vec(qd, q) < 1e-9
    m.close()


def test_error_behaviour_matches_reference():
    d = make_glm_data("bernoulli_logit", 1000, 3)
    y = d["y"].copy()
    y[17] = 2
    m = GLMModel("bernoulli_logit", d["X"], y)
    with pytest.raises(stan_b200.DomainError):        # check_bounded(y, 0, 1): domain_error on every call
        m.log_prob_grad(np.zeros(m.P))
    m.close()
    d = make_glm_data("poisson_log", 1000, 3)
    m = GLMModel("poisson_log", d["X"], d["y"])
    with pytest.raises(stan_b200.DomainError):
        m.log_prob_grad(np.array([np.nan, 0, 0, 0]))
    with pytest.raises(stan_b200.DomainError):        # exp overflow -> non-finite -> domain_error
        m.log_prob_grad(np.array([800.0, 0, 0, 0]))
    with pytest.raises(stan_b200.InvalidArgument):    # wrong theta length
        m.log_prob_grad(np.zeros(m.P + 1))
    lp, g = m.log_prob_grad(np.zeros(m.P))           # the handle is still usable afterwards
    assert np.isfinite(lp)
    # leapfrog into a domain error: V=+inf, g negated (base_hamiltonian.hpp:65-69)
    q = np.zeros(m.P)
    p = np.array([1e6, 0, 0, 0.0])
    m.set_state(q, p, -g, -lp)
    q1, p1, g1, V1 = m.leapfrog(1.0)
    assert V1 == np.inf and np.array_equal(g1, g)
    m.close()


@pytest.mark.parametrize("fam,K", [("bernoulli_logit", 6), ("poisson_log", 6), ("normal_id", 6), ("bernoulli_logit", 300)])
def test_nan_in_x_is_a_domain_error_in_every_kernel(fam, K):
    """A NaN in the matrix of independent variables makes the row's term NaN, the sum non-finite, and the reference
    throws domain_error from check_finite("Matrix of independent variables", ...) (bernoulli_logit_glm_lpmf.hpp:128-131,
    poisson_log_glm_lpmf.hpp:120-123, normal_id_glm_lpdf.hpp:192-197; rev test `..._glm_error_checking`, xw3) -- also when
    the weight of that column is zero (0 * NaN).  Narrow / wide kernel, and per lane in the batched kernels."""
    d = make_glm_data(fam, 5_000, K)
    X = d["X"].copy()
    X[4_321, 2] = np.nan
    m = GLMModel(fam, X, d["y"])
    for th in (0.1 * np.ones(m.P), np.zeros(m.P)):
        with pytest.raises(stan_b200.DomainError):
            m.log_prob_grad(th)
    if K <= 208:
        m.batch_reserve(20)
        th = 0.1 * np.random.default_rng(0).standard_normal((20, m.P))
        for n in (3, 12, 20):                      # few-chain kernel, row-split DMMA, DMMA
            assert m.log_prob_grad_batched(th[:n])[2].all()
    m.close()


def test_empty_and_tiny():
    m = GLMModel("bernoulli_logit", np.zeros((0, 3)), np.zeros(0, np.int32))
    from oracle.oracle import PortOracle
    po = PortOracle("bernoulli_logit", np.zeros((0, 3)), np.zeros(0, np.int32))
    th = 0.1 * np.arange(1, m.P + 1)
    lp, g = m.log_prob_grad(th)
    lp_r, g_r = po.log_prob_grad(th)
    assert rel_err(lp, lp_r) < TOL and rel_err_vec(g, g_r) < TOL
    m.close()


def test_slots_are_independent():
    d = make_glm_data("normal_id", 30_000, 16)
    m = GLMModel("normal_id", d["X"], d["y"], n_slots=3)
    ths = theta_points(m.P, n_random=2)
    ref = [m.log_prob_grad(th, slot=0) for th in ths]
    for s in range(3):
        lp, g = m.log_prob_grad(ths[s], slot=s)
        assert lp == ref[s][0] and np.array_equal(g, ref[s][1])
    m.close()


def test_device_resident_data_path():
    """data_on_device=1: X generated on the GPU (torch is only plumbing here)."""
    import torch
    N, K = 100_000, 24
    gen = torch.Generator(device="cuda").manual_seed(7)
    Xd = torch.randn((K, N), generator=gen, device="cuda", dtype=torch.float64)   # column-major N x K
    yd = (torch.rand(N, generator=gen, device="cuda") < 0.4).to(torch.int32)
    m = GLMModel("bernoulli_logit", Xd.data_ptr(), yd.data_ptr(), data_on_device=True, N=N, K=K, ldx=N)
    X = Xd.cpu().numpy().T
    from oracle.oracle import PortOracle
    po = PortOracle("bernoulli_logit", X, yd.cpu().numpy())
    th = theta_points(m.P)[1]
    lp, g = m.log_prob_grad(th)
    lp_r, g_r = po.log_prob_grad(th)
    assert rel_err(lp, lp_r) < TOL and rel_err_vec(g, g_r) < TOL
    m.close()


def test_full_size_config2_against_reference():
    """BASELINE configs[1] at full size: N=10M, K=100 (X = 8 GB, generated on the GPU), compared with
    the reference itself evaluated on the host copy of the very same buffers, plus the additive
    'checksum of checksums' property: the likelihood part of ten 1M-row shards sums to the whole."""
    import torch
    from oracle.oracle import PortOracle, RefOracle
    from stan_b200.synth import make_logistic_shard
    if not RefOracle.available():
        pytest.skip("oracle/_ref not present")
    N, K = 10_000_000, 100
    dev = torch.device("cuda", 0)
    X, y, _, _ = make_logistic_shard(torch, dev, N, K)
    m = GLMModel("bernoulli_logit", X.data_ptr(), y.data_ptr(), data_on_device=True, N=N, K=K, ldx=N)
    th = 0.05 * np.random.default_rng(11).standard_normal(K + 1)
    lp, g = m.log_prob_grad(th)
    # shards (device-resident views of the same buffers: column-major with ldx = N)
    prior = PortOracle("bernoulli_logit", np.zeros((0, K)), np.zeros(0, np.int32))
    lp0, g0 = prior.log_prob_grad(th)
    lp_sum, g_sum = lp0, g0.copy()
    for s in range(10):
        r0 = s * 1_000_000
        ms = GLMModel("bernoulli_logit", X.data_ptr() + 8 * r0, y.data_ptr() + 4 * r0, data_on_device=True,
                      N=1_000_000, K=K, ldx=N)
        lps, gs = ms.log_prob_grad(th)
        lp_sum += lps - lp0
        g_sum += gs - g0
        ms.close()
    assert rel_err(lp, lp_sum) < 1e-12 and rel_err_vec(g, g_sum) < 1e-12
    Xh, yh = X.cpu().numpy().T, y.cpu().numpy()
    del X, y
    torch.cuda.empty_cache()
    ro = RefOracle("bernoulli_logit", Xh, yh)
    lp_r, g_r = ro.log_prob_grad(th)
    assert rel_err(lp, lp_r) < TOL, (lp, lp_r)
    assert rel_err_vec(g, g_r) < TOL
    m.close()


# ---------------------------------------------------------------------------------------------
# Fused group path (SURVEY 2.3 contract for a16: 1 read of X, 0 N-vector traffic, 1 launch): rows are sorted by
# group, every CTA owns a contiguous panel range and accumulates its few groups' residual sums on chip.
# ---------------------------------------------------------------------------------------------
GROUP_SHAPES = [
    # family, N, K, G, expect_fused
    ("poisson_log", 80_000, 50, 1000, True),
    ("poisson_log", 5_000, 8, 3000, True),        # ~19 groups per panel: the boundary-panel path on every panel
    ("normal_id", 9_999, 31, 17, True),
    ("bernoulli_logit", 30, 3, 50, True),         # most groups have no rows at all
    ("neg_binomial_2_log", 20_000, 9, 37, True),
    ("binomial_logit", 12_345, 6, 5, True),
    ("poisson_log", 148 * 32 * 40, 4, 100_000, False),   # a CTA's range meets > 512 groups: unfused fallback
]


@pytest.mark.parametrize("fam,N,K,G,fused", GROUP_SHAPES)
def test_group_path_fused_matches_oracle_and_unfused(fam, N, K, G, fused, monkeypatch):
    d = make_glm_data(fam, N, K, G)
    kw = {"trials": d["trials"]} if fam == "binomial_logit" else {}
    orc = oracle_for(fam, d, G, **kw)
    m = GLMModel(fam, d["X"], d["y"], d["group"], G, **kw)
    monkeypatch.setenv("B200GLM_NO_GROUP_FUSION", "1")
    m_unfused = GLMModel(fam, d["X"], d["y"], d["group"], G, **kw)
    monkeypatch.delenv("B200GLM_NO_GROUP_FUSION")
    for th in theta_points(m.P, n_random=2, scale=0.1):
        n0 = m.launch_count()
        lp, g = m.log_prob_grad(th)
        assert m.launch_count() - n0 == (1 if fused else 3)     # one launch, no N-vector of residuals
        lp_r, g_r = orc.log_prob_grad(th)
        assert rel_err(lp, lp_r) < TOL, (lp, lp_r)
        assert rel_err_vec(g, g_r) < TOL
        lp_u, g_u = m_unfused.log_prob_grad(th)
        assert rel_err(lp, lp_u) < 1e-13 and rel_err_vec(g, g_u) < 1e-13
        assert rel_err(m.log_prob(th, False, True), orc.log_prob(th, False, True)) < TOL
    a, b = m.log_prob_grad(th), m.log_prob_grad(th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])          # deterministic
    # leapfrog steps stay on the device and follow the oracle's integrator
    from oracle.oracle import PortOracle
    po = PortOracle(fam, d["X"], d["y"], d["group"], G, **kw)
    rng = np.random.default_rng(3)
    th, p0 = 0.05 * rng.standard_normal(m.P), rng.standard_normal(m.P)
    lp, g = m.log_prob_grad(th)
    m.set_state(th, p0, -g, -lp)
    q, p, gg, V = th, p0, -g, -lp
    for _ in range(3):
        q1, p1, g1, V1 = m.leapfrog(1e-3)
        q, p, gg, V = po.leapfrog(1e-3, np.ones(m.P), q, p, gg, V)
    assert np.max(np.abs(q1 - q)) < 1e-12 and rel_err(V1, V) < TOL and rel_err_vec(g1, gg) < TOL
    m.close()
    m_unfused.close()


@pytest.mark.parametrize("fam,N,K,flags,chunk", [
    ("bernoulli_logit", 10_003, 20, 0, 3_200), ("normal_id", 5_000, 300, 0, 1_600), ("binomial_logit", 4_097, 9, 0, 4_096),
    ("poisson_log", 2_000, 7, 1, 640), ("neg_binomial_2_log", 3_001, 5, 0, 3_200)])
def test_streamed_construction_equals_one_shot(fam, N, K, flags, chunk):
    """B200GLM_FLAG_STREAMED + b200glm_append_rows + b200glm_finalize (X never resident twice): bitwise the same model
    as b200glm_create from the whole matrix; host chunks and device chunks; errors for misuse."""
    import torch
    from stan_b200.model import InvalidArgument
    d = make_glm_data(fam, N, K)
    kw = {"trials": d["trials"]} if fam == "binomial_logit" else {}
    one = GLMModel(fam, d["X"], d["y"], flags=flags, **kw)
    th = theta_points(one.P, n_random=1, scale=0.1)[1]
    want = one.log_prob_grad(th), one.log_prob(th, False, True)
    for on_device in (False, True):
        m = GLMModel.streamed(fam, N, K, data_on_device=on_device, flags=flags)
        with pytest.raises(InvalidArgument):
            m.log_prob_grad(th)                                    # not finalized
        for r0 in range(0, N, chunk):
            r1 = min(N, r0 + chunk)
            Xc, yc = np.asfortranarray(d["X"][r0:r1]), d["y"][r0:r1]
            tc = d["trials"][r0:r1] if fam == "binomial_logit" else None
            if on_device:
                Xt = torch.from_numpy(np.ascontiguousarray(Xc.T)).cuda()       # (K, n) row-major == column-major n x K
                yt = torch.from_numpy(np.ascontiguousarray(yc)).cuda()
                tt = torch.from_numpy(np.ascontiguousarray(tc)).cuda() if tc is not None else None
                m.append_rows(Xt.data_ptr(), yt.data_ptr(), tt.data_ptr() if tt is not None else None, n=r1 - r0)
                torch.cuda.synchronize()
            else:
                m.append_rows(Xc, yc, tc)
        m.finalize()
        got = m.log_prob_grad(th), m.log_prob(th, False, True)
        assert got[0][0] == want[0][0] and np.array_equal(got[0][1], want[0][1]) and got[1] == want[1]
        with pytest.raises(InvalidArgument):
            m.append_rows(np.asfortranarray(d["X"][:32]), d["y"][:32], d["trials"][:32] if kw else None)   # already complete
        m.close()
    m = GLMModel.streamed(fam, N, K, data_on_device=False, flags=flags)
    with pytest.raises(InvalidArgument):
        m.append_rows(np.asfortranarray(d["X"][:33]), d["y"][:33], d["trials"][:33] if kw else None)       # not a multiple of the panel height
    with pytest.raises(InvalidArgument):
        m.finalize()                                               # rows missing
    m.close()
    one.close()


def test_launch_timeline_and_peak_measurements():
    """Measurement entry points of the C ABI: b200glm_timeline_* (per-phase %globaltimer stamps of a gradient launch,
    monotone along the launch, result unchanged with the stamps on) and b200glm_measure_peaks (read-only-stream HBM
    bandwidth and fp64 DMMA peak in the caller's process: plausible for a B200)."""
    from stan_b200 import _capi
    d = make_glm_data("bernoulli_logit", 200_000, 40)
    m = GLMModel("bernoulli_logit", d["X"], d["y"])
    th = theta_points(m.P, n_random=1, scale=0.1)[1]
    want = m.log_prob_grad(th)
    m.timeline_enable(True)
    got = m.log_prob_grad(th)
    assert got[0] == want[0] and np.array_equal(got[1], want[1])
    T = m.timeline_read().astype(np.int64)
    grid = T.shape[0] - 2
    cta, tail, fine = T[:grid, :6], T[grid], T[grid + 1]
    assert np.all(np.diff(cta, axis=1) >= 0) and np.all(cta[:, 0] > 0)       # entry <= ... <= ticket, every CTA
    assert 0 <= tail[7] < grid and tail[0] >= cta[:, 5].max() - 64           # the last CTA's ticket comes last (32 ns ticks)
    assert tail[0] <= fine[0] <= fine[1] <= fine[2] <= tail[1] <= tail[2] <= fine[3] <= fine[4] <= tail[3]
    assert (tail[3] - cta[:, 0].min()) < 5_000_000                           # a 64 MB launch ends within 5 ms
    m.timeline_enable(False)
    with pytest.raises(stan_b200.InvalidArgument):
        m.timeline_read()
    m.close()
    read_gbs, dmma = _capi.measure_peaks(0, read=True, dmma=True)
    assert 4000 < read_gbs < 9000 and 25 < dmma < 45
