"""GPU suite: the device path of the reference's two class-outcome GLMs, ordered_logistic_glm_lpmf and
categorical_logit_glm_lpmf (glm_class_kernel.cuh), through the C ABI, against
  * the committed goldens generated from the compiled reference (tests/golden/glm_class_models_golden.json),
  * the CPU checker live (compiled reference when present, else the C port) on shapes that exercise every kernel
    instantiation, ragged panels, the leapfrog tail, error behaviour and the reference's early-return quirks.
Bar: 1e-10 relative (north_star); gradient entries scaled as in conftest.rel_err_vec."""
import numpy as np
import pytest

from conftest import CLASS_GOLDEN_NAMES, rel_err, rel_err_vec, unhex
from stan_b200 import GLMModel, make_glm_data
from stan_b200.model import DomainError, InvalidArgument

pytestmark = pytest.mark.gpu
TOL = 1e-10


def checker(fam, d, C):
    from oracle.oracle import PortOracle, RefOracle
    cls = RefOracle if RefOracle.available() else PortOracle
    return cls(fam, d["X"], d["y"], n_classes=C)


@pytest.mark.parametrize("name", CLASS_GOLDEN_NAMES)
def test_cuda_matches_reference_golden(golden_classes, name):
    c = golden_classes[name]
    m = GLMModel(c["family"], c["X"], c["y"], n_classes=c["n_classes"])
    assert m.num_params_r() == len(c["evals"][0]["theta"])
    for e in c["evals"]:
        th = unhex(e["theta"])
        for key, ref in e["lp_grad"].items():
            lp, g = m.log_prob_grad(th, int(key[0]), int(key[1]))
            assert rel_err(lp, float.fromhex(ref["lp"])) < TOL, (name, key)
            assert rel_err_vec(g, unhex(ref["grad"])) < TOL, (name, key)
        for key, ref in e["lp_double"].items():
            assert rel_err(m.log_prob(th, int(key[0]), int(key[1])), float.fromhex(ref)) < TOL, (name, key)
    lf = c["leapfrog"]
    m.set_state(unhex(lf["q0"]), unhex(lf["p0"]), unhex(lf["g0"]), float.fromhex(lf["V0"]))
    q, p, g, V = m.leapfrog(lf["eps"], unhex(lf["inv_metric"]))
    assert rel_err_vec(q, unhex(lf["q1"])) < TOL and rel_err_vec(p, unhex(lf["p1"])) < TOL
    assert rel_err_vec(g, unhex(lf["g1"])) < TOL and rel_err(V, float.fromhex(lf["V1"])) < TOL
    m.close()


SHAPES = [
    # family, N, K, classes -- one per kernel instantiation and its boundaries, ragged N, K = 0
    ("ordered_logistic", 20_000, 20, 5), ("ordered_logistic", 4_097, 32, 3), ("ordered_logistic", 9_999, 100, 9),
    ("ordered_logistic", 3_000, 104, 2), ("ordered_logistic", 2_000, 200, 16), ("ordered_logistic", 1_000, 256, 4),
    ("ordered_logistic", 777, 0, 6), ("ordered_logistic", 50_001, 3, 9),
    ("categorical_logit", 15_000, 12, 4), ("categorical_logit", 7_001, 30, 7), ("categorical_logit", 5_000, 104, 3),
    ("categorical_logit", 3_000, 200, 2), ("categorical_logit", 2_500, 56, 8), ("categorical_logit", 4_000, 16, 16),
    ("categorical_logit", 999, 0, 5), ("categorical_logit", 60_000, 100, 4),
]


@pytest.mark.parametrize("fam,N,K,C", SHAPES)
def test_cuda_matches_checker_live(fam, N, K, C):
    d = make_glm_data(fam, N, K, n_classes=C)
    orc = checker(fam, d, C)
    m = GLMModel(fam, d["X"], d["y"], n_classes=C)
    assert m.num_params_r() == orc.P
    rng = np.random.default_rng(17)
    for sc in (0.0, 0.2, 1.0):
        th = sc * rng.standard_normal(m.P)
        n0 = m.launch_count()
        lp, g = m.log_prob_grad(th)
        assert m.launch_count() - n0 == 1                     # one launch, epilogue fused
        lp_r, g_r = orc.log_prob_grad(th)
        assert rel_err(lp, lp_r) < TOL, (lp, lp_r)
        assert rel_err_vec(g, g_r) < TOL
        lp0, g0 = m.log_prob_grad(th, False, True)
        lp0_r, g0_r = orc.log_prob_grad(th, False, True)
        assert rel_err(lp0, lp0_r) < TOL and rel_err_vec(g0, g0_r) < TOL
        assert rel_err(m.log_prob(th, False, True), orc.log_prob(th, False, True)) < TOL
        assert m.log_prob(th, True, True) == orc.log_prob(th, True, True)      # doubles under propto: Jacobian only
    a, b = m.log_prob_grad(th), m.log_prob_grad(th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])        # deterministic
    m.close()


@pytest.mark.parametrize("fam,K,C", [("ordered_logistic", 9, 5), ("categorical_logit", 6, 3)])
def test_device_resident_trajectory_follows_the_oracle(fam, K, C):
    from oracle.oracle import PortOracle
    d = make_glm_data(fam, 5_000, K, n_classes=C)
    po = PortOracle(fam, d["X"], d["y"], n_classes=C)
    m = GLMModel(fam, d["X"], d["y"], n_classes=C)
    rng = np.random.default_rng(3)
    th, p0, im = 0.1 * rng.standard_normal(m.P), rng.standard_normal(m.P), np.exp(0.3 * rng.standard_normal(m.P))
    lp, g = m.log_prob_grad(th)
    m.set_state(th, p0, -g, -lp)
    q, p, gg, V = th, p0, -g, -lp
    for i in range(10):
        q1, p1, g1, V1 = m.leapfrog(2e-3, im if i == 0 else None)
        q, p, gg, V = po.leapfrog(2e-3, im, q, p, gg, V)
    assert np.max(np.abs(q1 - q)) < 1e-12 and np.max(np.abs(p1 - p)) < 1e-9
    assert rel_err(V1, V) < TOL and rel_err_vec(g1, gg) < TOL
    m.close()


def test_error_behaviour_and_reference_quirks():
    d = make_glm_data("ordered_logistic", 300, 2, n_classes=3)
    y = d["y"].copy()
    y[4] = 4                                              # check_bounded(y, 1, N_classes), ordered_logistic_glm_lpmf.hpp:84
    m = GLMModel("ordered_logistic", d["X"], y, n_classes=3)
    with pytest.raises(DomainError):
        m.log_prob_grad(np.zeros(4))
    with pytest.raises(DomainError):
        m.log_prob(np.zeros(4), True, True)               # checked before include_summand
    m.close()
    m = GLMModel("ordered_logistic", d["X"], d["y"], n_classes=3)
    with pytest.raises(DomainError):                      # exp(u) underflows: cut-points not strictly increasing (:85)
        m.log_prob_grad(np.array([0.0, 0.0, 0.3, -800.0]))
    m.close()
    # one class: no cut-points, size_zero(cuts) => the likelihood is 0, the priors remain (:93-95)
    m = GLMModel("ordered_logistic", d["X"], np.ones(300, np.int32), n_classes=1)
    lp, g = m.log_prob_grad(np.array([0.5, -0.5]))
    assert rel_err(lp, -0.5 * 2 * (0.5 / 2.5) ** 2) < 1e-15
    m.close()
    d = make_glm_data("categorical_logit", 300, 2, n_classes=3)
    y = d["y"].copy()
    y[0] = 0
    m = GLMModel("categorical_logit", d["X"], y, n_classes=3)
    with pytest.raises(DomainError):                      # categorical outcome out of support (categorical...:74)
        m.log_prob_grad(np.zeros(9))
    m.close()
    # N_classes == 1 returns 0 BEFORE the bounds check (:70-72): a bad y is not an error there
    m = GLMModel("categorical_logit", d["X"], y, n_classes=1)
    lp, g = m.log_prob_grad(np.zeros(3))
    assert lp == 0.0 and not g.any()
    m.close()
    with pytest.raises(InvalidArgument):
        GLMModel("categorical_logit", d["X"], d["y"], n_classes=17)
    with pytest.raises(InvalidArgument):                  # K x classes beyond the register-resident shapes
        GLMModel("categorical_logit", np.zeros((64, 120)), np.ones(64, np.int32), n_classes=8)
    m = GLMModel("ordered_logistic", d["X"], d["y"], n_classes=3)
    with pytest.raises(InvalidArgument):
        m.glm_lpmf(0.0, np.zeros(2))                      # function-level entry serves families 0-4
    m.close()


def test_full_hbm_size_shard_additivity():
    """At a size the CPU checker cannot visit (2M rows x K = 100, 4 classes): the likelihood part is additive over row
    blocks -- two halves evaluated as separate handles sum to the whole -- and matches the checker on a 50k-row block."""
    import torch
    fam, N, K, C = "categorical_logit", 2_000_000, 100, 4
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn((K, N), generator=g, device="cuda", dtype=torch.float64)
    y = torch.randint(1, C + 1, (N,), generator=g, device="cuda", dtype=torch.int32)
    th = 0.05 * np.random.default_rng(2).standard_normal(C * (1 + K))

    def model(r0, r1):
        Xs = X[:, r0:r1].contiguous()
        ys = y[r0:r1].contiguous()
        m = GLMModel(fam, Xs.data_ptr(), ys.data_ptr(), data_on_device=True, N=r1 - r0, K=K, ldx=r1 - r0, n_classes=C)
        return m
    whole, a, b = model(0, N), model(0, N // 2 + 7), model(N // 2 + 7, N)
    pri = GLMModel(fam, np.zeros((0, K)), np.zeros(0, np.int32), n_classes=C)      # priors alone
    lw, gw = whole.log_prob_grad(th)
    la, ga = a.log_prob_grad(th)
    lb, gb = b.log_prob_grad(th)
    l0, g0 = pri.log_prob_grad(th)
    assert rel_err(lw - l0, (la - l0) + (lb - l0)) < 1e-12
    assert rel_err_vec(gw - g0, (ga - g0) + (gb - g0)) < 1e-12
    blk = model(0, 50_000)
    from oracle.oracle import PortOracle
    po = PortOracle(fam, X[:, :50_000].cpu().numpy().T, y[:50_000].cpu().numpy(), n_classes=C)
    lk, gk = blk.log_prob_grad(th)
    lr, gr = po.log_prob_grad(th)
    assert rel_err(lk, lr) < TOL and rel_err_vec(gk, gr) < TOL
    for m in (whole, a, b, pri, blk):
        m.close()
