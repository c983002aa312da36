"""GPU suite for the batched-chains path (fp64 DMMA GEMM pair, BASELINE configs[2]): every lane of a
batched call must equal the single-chain reference evaluation of that chain (1e-10 relative)."""
import numpy as np
import pytest

from conftest import rel_err, rel_err_vec
import stan_b200
from stan_b200 import GLMModel, make_glm_data

pytestmark = pytest.mark.gpu
TOL = 1e-10


def oracle_for(fam, d):
    from oracle.oracle import PortOracle, RefOracle
    cls = RefOracle if RefOracle.available() else PortOracle
    return cls(fam, d["X"], d["y"])


SHAPES = [
    # family, N, K, chains
    ("normal_id", 20_000, 200, 70),         # config 3's K; two chain blocks, the second one ragged
    ("normal_id", 3_001, 16, 5),
    ("bernoulli_logit", 10_000, 20, 64),    # config 1's shape
    ("bernoulli_logit", 5_000, 101, 33),    # K % 4 == 1: y column next to the last k4 step
    ("poisson_log", 8_000, 50, 130),
    ("bernoulli_logit", 4_097, 13, 3),
    ("normal_id", 2_000, 208, 9),           # largest supported K
    ("poisson_log", 1_000, 7, 200),
    ("bernoulli_logit", 31, 3, 2),          # a single partial panel
]


@pytest.mark.parametrize("fam,N,K,nch", SHAPES)
def test_batched_matches_oracle_per_chain(fam, N, K, nch):
    d = make_glm_data(fam, N, K)
    orc = oracle_for(fam, d)
    m = GLMModel(fam, d["X"], d["y"])
    m.batch_reserve(nch)
    rng = np.random.default_rng(3)
    th = 0.1 * rng.standard_normal((nch, m.P))
    th[0] = 0.0
    for propto, jac in ((1, 1), (0, 1), (1, 0)):
        lp, g, st = m.log_prob_grad_batched(th, propto, jac)
        assert not st.any()
        for c in range(nch) if nch <= 8 else (0, 1, nch // 2, nch - 1):
            lp_r, g_r = orc.log_prob_grad(th[c], propto, jac)
            assert rel_err(lp[c], lp_r) < TOL, (c, lp[c], lp_r)
            assert rel_err_vec(g[c], g_r) < TOL, c
    # a batched lane == the single-chain kernel on the same handle, to rounding
    lp1, g1 = m.log_prob_grad(th[1])
    lp, g, st = m.log_prob_grad_batched(th)
    assert rel_err(lp[1], lp1) < 1e-12 and rel_err_vec(g[1], g1) < 1e-12
    # deterministic, and independent of which lane / how many lanes a chain is evaluated in
    lp2, g2, _ = m.log_prob_grad_batched(th)
    assert np.array_equal(lp, lp2) and np.array_equal(g, g2)
    lp3, g3, _ = m.log_prob_grad_batched(th[::-1].copy())
    assert np.array_equal(lp3[::-1], lp) and np.array_equal(g3[::-1], g)
    m.close()


def test_batched_leapfrog_matches_oracle():
    """Lanes advance different chain slots with different step sizes and metrics; 6 steps, state resident."""
    from oracle.oracle import PortOracle
    fam, N, K, nch = "normal_id", 6_000, 24, 37
    d = make_glm_data(fam, N, K)
    po = PortOracle(fam, d["X"], d["y"])
    m = GLMModel(fam, d["X"], d["y"])
    m.batch_reserve(64)
    rng = np.random.default_rng(9)
    q = 0.05 * rng.standard_normal((nch, m.P))
    p = rng.standard_normal((nch, m.P))
    im = np.exp(0.3 * rng.standard_normal((nch, m.P)))
    eps = 1e-3 * (1 + rng.random(nch)) * np.where(rng.random(nch) < 0.3, -1.0, 1.0)
    chains = rng.permutation(64)[:nch].astype(np.int32)
    lp, g, _ = m.log_prob_grad_batched(q)
    g, V = -g, -lp
    m.set_state_batched(q, p, g, V, im, chains)
    ref = [(q[i], p[i], g[i], V[i]) for i in range(nch)]
    for step in range(6):
        qd, pd, gd, Vd, st = m.leapfrog_batched(eps, chains)
        assert not st.any()
        ref = [po.leapfrog(eps[i], im[i], *ref[i]) for i in range(nch)]
    for i in range(nch):
        assert rel_err_vec(qd[i], ref[i][0]) < 1e-9 and rel_err_vec(pd[i], ref[i][1]) < 1e-9
        assert rel_err_vec(gd[i], ref[i][2]) < 1e-9 and rel_err(Vd[i], ref[i][3]) < 1e-9
    # a subset of the slots, in another order: the others keep their state
    sub = np.array([chains[5], chains[2]], dtype=np.int32)
    q2, p2, g2, V2, _ = m.leapfrog_batched(np.array([eps[5], eps[2]]), sub)
    r5, r2 = po.leapfrog(eps[5], im[5], *ref[5]), po.leapfrog(eps[2], im[2], *ref[2])
    assert rel_err_vec(q2[0], r5[0]) < 1e-9 and rel_err_vec(q2[1], r2[0]) < 1e-9
    m.close()


def test_batched_domain_error_is_per_chain():
    d = make_glm_data("poisson_log", 2_000, 4)
    m = GLMModel("poisson_log", d["X"], d["y"])
    m.batch_reserve(8)
    th = np.zeros((3, m.P))
    th[1, 0] = 800.0                      # exp overflow in chain 1 only
    lp, g, st = m.log_prob_grad_batched(th)
    assert list(st) == [0, 1, 0]
    lp0, g0 = m.log_prob_grad(th[0])
    assert rel_err(lp[0], lp0) < 1e-12 and rel_err(lp[2], lp0) < 1e-12
    # leapfrog into a domain error: V = +inf, g negated, the other lanes unaffected
    q = np.zeros((2, m.P))
    p = np.zeros((2, m.P))
    p[1, 0] = 1e6
    gg = np.stack([-g0, -g0])
    m.set_state_batched(q, p, gg, np.array([-lp0, -lp0]))
    q1, p1, g1, V1, st = m.leapfrog_batched(np.array([1e-3, 1.0]))
    assert list(st) == [0, 1] and V1[1] == np.inf and np.array_equal(g1[1], g0) and np.isfinite(V1[0])
    m.close()


def test_batched_argument_errors():
    d = make_glm_data("bernoulli_logit", 500, 3)
    m = GLMModel("bernoulli_logit", d["X"], d["y"])
    with pytest.raises(stan_b200.InvalidArgument):      # reserve first
        m.log_prob_grad_batched(np.zeros((2, m.P)))
    m.batch_reserve(4)
    with pytest.raises(stan_b200.InvalidArgument):      # more lanes than reserved
        m.log_prob_grad_batched(np.zeros((5, m.P)))
    with pytest.raises(stan_b200.InvalidArgument):      # chain slot out of range
        m.leapfrog_batched(np.array([0.1]), np.array([4], dtype=np.int32))
    m.close()
    d = make_glm_data("bernoulli_logit", 500, 300)
    m = GLMModel("bernoulli_logit", d["X"], d["y"])
    with pytest.raises(stan_b200.InvalidArgument):      # K > 208
        m.batch_reserve(4)
    m.close()


@pytest.mark.parametrize("fam,N,K", [("normal_id", 20_000, 200), ("bernoulli_logit", 9_001, 100), ("poisson_log", 4_000, 17),
                                     ("normal_id", 150, 8)])
def test_row_split_small_batches(fam, N, K):
    """Batches of <= 16 lanes run the row-split variant of the DMMA kernel (every warp pair on its own row panels
    of the same 16 chains, partial blocks folded by the reduce kernel): same answers as the oracle, as the same
    chains evaluated inside a larger (normal-mode) batch to rounding, deterministic, leapfrog included."""
    from oracle.oracle import PortOracle
    d = make_glm_data(fam, N, K)
    po = PortOracle(fam, d["X"], d["y"])
    m = GLMModel(fam, d["X"], d["y"])
    m.batch_reserve(64)
    rng = np.random.default_rng(21)
    th = 0.1 * rng.standard_normal((40, m.P))
    lp_big, g_big, _ = m.log_prob_grad_batched(th)              # 40 lanes: normal mode
    for n in (1, 5, 16):
        lp, g, st = m.log_prob_grad_batched(th[:n])             # row-split mode
        assert not st.any()
        for c in range(n):
            lp_r, g_r = po.log_prob_grad(th[c])
            assert rel_err(lp[c], lp_r) < TOL and rel_err_vec(g[c], g_r) < TOL, (n, c)
            assert rel_err(lp[c], lp_big[c]) < 1e-12 and rel_err_vec(g[c], g_big[c]) < 1e-12
        lp2, g2, _ = m.log_prob_grad_batched(th[:n])
        assert np.array_equal(lp, lp2) and np.array_equal(g, g2)
    # three leapfrog steps of 7 chain slots scattered over the 64
    chains = rng.permutation(64)[:7].astype(np.int32)
    q, p = th[:7].copy(), rng.standard_normal((7, m.P))
    im = np.exp(0.3 * rng.standard_normal((7, m.P)))
    lp, g, _ = m.log_prob_grad_batched(q)
    m.set_state_batched(q, p, -g, -lp, im, chains)
    eps = 1e-3 * (1 + rng.random(7))
    ref = [(q[i], p[i], -g[i], -lp[i]) for i in range(7)]
    for _ in range(3):
        qd, pd, gd, Vd, st = m.leapfrog_batched(eps, chains)
        ref = [po.leapfrog(eps[i], im[i], *ref[i]) for i in range(7)]
    for i in range(7):
        assert rel_err_vec(qd[i], ref[i][0]) < 1e-9 and rel_err_vec(gd[i], ref[i][2]) < 1e-9
        assert rel_err(Vd[i], ref[i][3]) < 1e-9
    m.close()


def test_batch_workspace_grows_on_a_larger_reservation():
    """A second b200glm_batch_reserve with MORE chains re-reserves the workspace (round 1 returned INVALID, which the
    batched driver read as "shape not supported" and silently fell back to lane-by-lane launches)."""
    from oracle.oracle import PortOracle
    d = make_glm_data("poisson_log", 3_000, 9)
    po = PortOracle("poisson_log", d["X"], d["y"])
    m = GLMModel("poisson_log", d["X"], d["y"])
    th = 0.1 * np.random.default_rng(4).standard_normal((70, m.P))
    m.batch_reserve(8)
    lp8, g8, _ = m.log_prob_grad_batched(th[:8])
    m.batch_reserve(70)                                    # grows
    lp, g, st = m.log_prob_grad_batched(th)
    assert not st.any()
    for c in range(8):      # 8 lanes ran the row-split variant, 70 the normal mode: equal to rounding
        assert rel_err(lp[c], lp8[c]) < 1e-12 and rel_err_vec(g[c], g8[c]) < 1e-12
    for c in (0, 9, 69):
        lp_r, g_r = po.log_prob_grad(th[c])
        assert rel_err(lp[c], lp_r) < TOL and rel_err_vec(g[c], g_r) < TOL
    m.batch_reserve(16)                                    # smaller: keeps the larger workspace
    assert m.log_prob_grad_batched(th)[2].sum() == 0
    m.close()


@pytest.mark.parametrize("fam,N,K", [("bernoulli_logit", 9_001, 100), ("normal_id", 6_000, 128), ("poisson_log", 4_000, 17),
                                     ("bernoulli_logit", 70_000, 33), ("normal_id", 150, 8), ("poisson_log", 31, 3)])
def test_few_chain_fma_kernel(fam, N, K, monkeypatch):
    """Batches of <= 4 lanes with K <= 128 run glm_multi_kernel (four chains per pass on the FMA path, X read from
    shared memory once per phase for all of them): same answers as the oracle, as the DMMA path evaluating the same chains
    (B200GLM_NO_MULTI=1) to rounding, independent of the lane a chain sits in and of how many lanes there are, and through
    the leapfrog entry point."""
    from oracle.oracle import PortOracle
    d = make_glm_data(fam, N, K)
    po = PortOracle(fam, d["X"], d["y"])
    m = GLMModel(fam, d["X"], d["y"])
    m.batch_reserve(8)
    monkeypatch.setenv("B200GLM_NO_MULTI", "1")
    m_dmma = GLMModel(fam, d["X"], d["y"])
    m_dmma.batch_reserve(8)
    monkeypatch.delenv("B200GLM_NO_MULTI")
    rng = np.random.default_rng(8)
    th = 0.1 * rng.standard_normal((4, m.P))
    lp4, g4, st = m.log_prob_grad_batched(th)
    assert not st.any()
    lpd, gd, _ = m_dmma.log_prob_grad_batched(th)
    for c in range(4):
        lp_r, g_r = po.log_prob_grad(th[c])
        assert rel_err(lp4[c], lp_r) < TOL and rel_err_vec(g4[c], g_r) < TOL, c
        assert rel_err(lp4[c], lpd[c]) < 1e-12 and rel_err_vec(g4[c], gd[c]) < 1e-12, c
    for n in (1, 2, 3):                                   # fewer lanes: the same chains, bit for bit
        lp, g, _ = m.log_prob_grad_batched(th[:n])
        assert np.array_equal(lp, lp4[:n]) and np.array_equal(g, g4[:n])
    th7 = np.concatenate([th, 0.1 * rng.standard_normal((3, m.P))])   # 5-8 lanes: two passes of four chains
    lp7, g7, st7 = m.log_prob_grad_batched(th7)
    assert not st7.any() and np.array_equal(lp7[:4], lp4) and np.array_equal(g7[:4], g4)
    lpd7, gd7, _ = m_dmma.log_prob_grad_batched(th7)
    for c in range(4, 7):
        lp_r, g_r = po.log_prob_grad(th7[c])
        assert rel_err(lp7[c], lp_r) < TOL and rel_err_vec(g7[c], g_r) < TOL, c
        assert rel_err(lp7[c], lpd7[c]) < 1e-12 and rel_err_vec(g7[c], gd7[c]) < 1e-12, c
    lp_r4, g_r4, _ = m.log_prob_grad_batched(th[::-1].copy())
    assert np.array_equal(lp_r4[::-1], lp4) and np.array_equal(g_r4[::-1], g4)
    for propto, jac in ((0, 1), (1, 0)):
        lp, g, _ = m.log_prob_grad_batched(th[:3], propto, jac)
        for c in range(3):
            lp_r, g_r = po.log_prob_grad(th[c], propto, jac)
            assert rel_err(lp[c], lp_r) < TOL and rel_err_vec(g[c], g_r) < TOL
    # three leapfrog steps of 3 chain slots scattered over the 8
    chains = np.array([6, 1, 4], dtype=np.int32)
    q, p = th[:3].copy(), rng.standard_normal((3, m.P))
    im = np.exp(0.3 * rng.standard_normal((3, m.P)))
    lp, g, _ = m.log_prob_grad_batched(q)
    m.set_state_batched(q, p, -g, -lp, im, chains)
    eps = 1e-3 * (1 + rng.random(3))
    ref = [(q[i], p[i], -g[i], -lp[i]) for i in range(3)]
    for _ in range(3):
        qd, pd, gd2, Vd, st = m.leapfrog_batched(eps, chains)
        ref = [po.leapfrog(eps[i], im[i], *ref[i]) for i in range(3)]
    for i in range(3):
        assert rel_err_vec(qd[i], ref[i][0]) < 1e-9 and rel_err_vec(gd2[i], ref[i][2]) < 1e-9
        assert rel_err(Vd[i], ref[i][3]) < 1e-9
    m.close()
    m_dmma.close()


@pytest.mark.parametrize("N,K", [(9_001, 100), (4_000, 17), (2_500, 22), (31, 3)])
def test_binomial_logit_in_the_batched_kernels(N, K):
    """binomial_logit_glm (population sizes in panel column K + 1) through all three batched kernels: the few-chain FMA
    kernel (<= 8 lanes; branch-free link_bf<>), the row-split DMMA variant (9-16 lanes) and the DMMA kernel (40 lanes;
    link_ext<>): per chain against the oracle, the kernels against each other and against the single-chain kernel to
    rounding, propto on / off (the binomial-coefficient constant), and three leapfrog steps."""
    from oracle.oracle import PortOracle
    fam = "binomial_logit"
    d = make_glm_data(fam, N, K)
    po = PortOracle(fam, d["X"], d["y"], trials=d["trials"])
    m = GLMModel(fam, d["X"], d["y"], trials=d["trials"])
    m.batch_reserve(64)
    rng = np.random.default_rng(17)
    th = 0.1 * rng.standard_normal((40, m.P))
    big = {}
    for propto, jac in ((1, 1), (0, 1)):
        lp_big, g_big, st = m.log_prob_grad_batched(th, propto, jac)       # 40 lanes: DMMA kernel
        assert not st.any()
        big[propto] = (lp_big, g_big)
        for c in (0, 1, 17, 39):
            lp_r, g_r = po.log_prob_grad(th[c], propto, jac)
            assert rel_err(lp_big[c], lp_r) < TOL and rel_err_vec(g_big[c], g_r) < TOL, (propto, c)
    lp_big, g_big = big[1]
    lp1, g1 = m.log_prob_grad(th[1])
    assert rel_err(lp_big[1], lp1) < 1e-12 and rel_err_vec(g_big[1], g1) < 1e-12
    for n in (1, 4, 7, 12):                                                 # few-chain kernel (<= 8), row-split (12)
        lp, g, st = m.log_prob_grad_batched(th[:n])
        assert not st.any()
        for c in range(n):
            lp_r, g_r = po.log_prob_grad(th[c])
            assert rel_err(lp[c], lp_r) < TOL and rel_err_vec(g[c], g_r) < TOL, (n, c)
            assert rel_err(lp[c], lp_big[c]) < 1e-12 and rel_err_vec(g[c], g_big[c]) < 1e-12, (n, c)
    q, p = th[:5].copy(), rng.standard_normal((5, m.P))
    lp, g, _ = m.log_prob_grad_batched(q)
    m.set_state_batched(q, p, -g, -lp)
    eps = 1e-3 * (1 + rng.random(5))
    ref = [(q[i], p[i], -g[i], -lp[i]) for i in range(5)]
    for _ in range(3):
        qd, pd, gd, Vd, st = m.leapfrog_batched(eps)
        ref = [po.leapfrog(eps[i], np.ones(m.P), *ref[i]) for i in range(5)]
    for i in range(5):
        assert rel_err_vec(qd[i], ref[i][0]) < 1e-9 and rel_err_vec(gd[i], ref[i][2]) < 1e-9
        assert rel_err(Vd[i], ref[i][3]) < 1e-9
    m.close()


def test_binomial_logit_device_nuts_matches_the_service():
    """Device-side NUTS on a binomial_logit model (few-chain kernel, 4 chains) == the unmodified service, same seeds."""
    from stan_b200 import stan_service
    if not stan_service.available():
        pytest.skip("libb200stan.so not built")
    d = make_glm_data("binomial_logit", 3_000, 6)
    m = stan_service.StanGLM("binomial_logit", d["X"], d["y"], trials=d["trials"])
    kw = dict(num_chains=4, seed=5, num_warmup=60, num_samples=30, delta=0.8)
    dev, seq = m.nuts_device(**kw), m.nuts(num_threads=2, **kw)
    m.close()
    a, b = dev["warmup_draws"][:, :5, :], seq["warmup_draws"][:, :5, :]
    assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])
    assert np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6
    assert np.abs(dev["draws"][:, :, 7:].mean(axis=(0, 1)) - seq["draws"][:, :, 7:].mean(axis=(0, 1))).max() < \
        1.0 * seq["draws"][:, :, 7:].std(axis=(0, 1)).max()
