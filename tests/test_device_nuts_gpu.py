"""GPU suite for the device-side NUTS transition + adaptation (stan_b200/csrc/nuts_tree.cuh,
nuts_kernels.cuh, b200glm_nuts_*, stan_b200/cpp/b200/device_nuts.hpp): against the reference's own unmodified service on
the same device model (same seeds => the same chains until rounding is amplified), against the compiled reference on the
CPU (posterior within Monte-Carlo standard error), and the reference's parallel-match property at 1024 chains.
The state machine's logic is also pinned without a GPU in tests/test_device_nuts_host.py."""
import ctypes as C

import numpy as np
import pytest

from stan_b200 import _capi, make_glm_data
from stan_b200 import stan_service

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not stan_service.available(), reason="libb200stan.so not built")]


def ref_oracle():
    from oracle.oracle import RefOracle
    if not RefOracle.available():
        pytest.skip("oracle/_ref not present")
    return RefOracle


@pytest.fixture(scope="module")
def device_pair():
    Ref = ref_oracle()
    d = make_glm_data("normal_id", 5_000, 12)
    m = stan_service.StanGLM("normal_id", d["X"], d["y"])
    kw = dict(num_chains=16, seed=99, num_warmup=300, num_samples=300, delta=0.8)
    dev = m.nuts_device(**kw)
    dev["counters"] = m.counters()
    seq = m.nuts(num_threads=4, **kw)
    m.close()
    ref = Ref("normal_id", d["X"], d["y"]).nuts(num_threads=4, **kw)
    return Ref, dev, seq, ref


def test_device_nuts_matches_the_reference_service_on_the_same_seeds(device_pair):
    """Chain i of the device-side sampler == chain i of stan::services::sample::hmc_nuts_diag_e_adapt (unmodified, host
    tree building) on the same device model: same tree depths / leapfrog counts / divergence flags and the same draws for
    the first iterations (the two differ in which kernel evaluates the gradient and in the summation order of the
    U-turn dot products, so rounding is amplified later on), and the same chains as the CPU reference."""
    Ref, dev, seq, ref = device_pair
    for other in (seq, ref):
        a, b = dev["warmup_draws"][:, :5, :], other["warmup_draws"][:, :5, :]
        assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])
        assert np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6
        assert np.max(np.abs(a[:, :, 0] - b[:, :, 0]) / np.abs(b[:, :, 0])) < 1e-9
        assert np.max(np.abs(a[:, :, 1:3] - b[:, :, 1:3])) < 1e-8      # accept_stat__, stepsize__ (init_stepsize)


def test_device_nuts_posterior_and_adaptation_within_mcse(device_pair):
    Ref, dev, seq, ref = device_pair
    P = dev["draws"].shape[2] - 7
    zs = []
    for k in range(P):
        a, b = dev["draws"][:, :, 7 + k].T, ref["draws"][:, :, 7 + k].T
        zs.append(abs(a.mean() - b.mean()) / np.hypot(Ref.mcse_mean(a), Ref.mcse_mean(b)))
        zs.append(abs(a.std(ddof=1) - b.std(ddof=1)) / np.hypot(Ref.mcse_sd(a), Ref.mcse_sd(b)))
        assert Ref.rhat(a) < 1.03
    assert max(zs) < 4.0, zs
    assert np.all(dev["draws"][:, :, 5] == 0)
    # the adapted step sizes and metrics are those of the reference's adaptation (same windows, same estimator)
    assert abs(np.median(dev["stepsize"]) / np.median(ref["stepsize"]) - 1.0) < 0.15
    assert abs(np.median(dev["inv_metric"] / ref["inv_metric"]) - 1.0) < 0.15
    # sampling draws carry the adapted step size
    assert np.allclose(dev["draws"][:, 0, 2], dev["stepsize"])


def test_device_nuts_rounds_and_traffic(device_pair):
    """One batched launch per round; the host supplies one vector of normal variates per momentum refresh and a few
    uniform variates per round; no per-leapfrog state comes back (the model's upload counter stays at the init calls)."""
    Ref, dev, seq, ref = device_pair
    n_eval = dev["draws"][:, :, 4].sum() + dev["warm_leapfrogs"].sum()
    assert dev["lanes"] >= n_eval
    assert dev["rounds"] < 0.25 * dev["lanes"]
    assert dev["normal_vectors"] >= 16 * 600


def test_device_nuts_parallel_match_at_1024_chains():
    """hmc_nuts_diag_e_adapt_parallel_match_test.cpp:73-140 at the chain count of BASELINE configs[2], with no host tree
    code at all: chains picked across the range equal the UNBATCHED reference service started with their chain ids."""
    d = make_glm_data("normal_id", 3_000, 4)
    m = stan_service.StanGLM("normal_id", d["X"], d["y"], n_slots=4)
    kw = dict(seed=2024, num_warmup=60, num_samples=40, delta=0.8)
    dev = m.nuts_device(num_chains=1024, **kw)
    assert np.all(np.isfinite(dev["draws"])) and dev["draws"].shape[0] == 1024
    assert dev["rounds"] < 0.02 * dev["lanes"]
    for c in (0, 1, 511, 777, 1023):
        one = m.nuts(num_chains=1, init_chain_id=1 + c, num_threads=1, **kw)
        a, b = dev["warmup_draws"][c, :5, :], one["warmup_draws"][0, :5, :]
        assert np.array_equal(a[:, 3:6], b[:, 3:6]), c
        assert np.max(np.abs(a[:, 7:] - b[:, 7:])) < 1e-6, c
    m.close()
    pooled = dev["draws"][:, :, 7:].reshape(-1, dev["draws"].shape[2] - 7)
    truth = np.concatenate([[d["truth"]["alpha"]], d["truth"]["beta"], [1.0]])
    assert np.all(np.abs(pooled.mean(axis=0) - truth) < 4.0 * pooled.std(axis=0))
    assert dev["draws"][:, :, 5].sum() == 0


@pytest.mark.parametrize("fam", ["bernoulli_logit", "poisson_log"])
def test_device_nuts_other_families_first_draws(fam):
    Ref = ref_oracle()
    d = make_glm_data(fam, 4_000, 8)
    m = stan_service.StanGLM(fam, d["X"], d["y"])
    kw = dict(num_chains=4, seed=5, num_warmup=100, num_samples=50, delta=0.8)
    dev = m.nuts_device(**kw)
    m.close()
    ref = Ref(fam, d["X"], d["y"]).nuts(num_threads=4, **kw)
    a, b = dev["warmup_draws"][:, :5, :], ref["warmup_draws"][:, :5, :]
    assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])
    assert np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6
    zs = []
    for k in range(dev["draws"].shape[2] - 7):
        x, y = np.concatenate([dev["warmup_draws"][:, 50:, 7 + k], dev["draws"][:, :, 7 + k]], axis=1).T, \
            np.concatenate([ref["warmup_draws"][:, 50:, 7 + k], ref["draws"][:, :, 7 + k]], axis=1).T
        zs.append(abs(x.mean() - y.mean()) / np.hypot(Ref.mcse_mean(x), Ref.mcse_mean(y)))
    assert max(zs) < 4.5, zs


def test_device_nuts_with_stepsize_jitter():
    """stepsize_jitter = 0.3 through all three drivers on the same seeds: one more uniform variate per transition, drawn
    before the momentum (base_hmc.hpp:195-200, base_nuts.hpp:80) -- the device-side chains stay the service's chains."""
    d = make_glm_data("bernoulli_logit", 3_000, 8)
    m = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"])
    kw = dict(num_chains=6, seed=21, num_warmup=60, num_samples=40, delta=0.8, stepsize_jitter=0.3)
    dev, seq, bat = m.nuts_device(**kw), m.nuts(num_threads=2, **kw), m.nuts_batched(**kw)
    plain = m.nuts_device(**dict(kw, stepsize_jitter=0.0))
    m.close()
    for other in (seq, bat):
        a, b = dev["warmup_draws"][:, :6, :], other["warmup_draws"][:, :6, :]
        assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])
        assert np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6
        assert np.max(np.abs(a[:, :, 1:3] - b[:, :, 1:3])) < 1e-8
    ratio = dev["draws"][:, :, 2] / dev["stepsize"][:, None]
    assert ratio.min() >= 0.7 - 1e-9 and ratio.max() <= 1.3 + 1e-9 and ratio.std() > 0.05
    assert np.all(plain["draws"][:, :, 2] == plain["stepsize"][:, None])
    assert not np.array_equal(plain["warmup_draws"][:, :6, 7:], dev["warmup_draws"][:, :6, 7:])


def test_nuts_c_abi_round_by_round():
    """b200glm_nuts_* called directly: reserve -> buffers -> init_chain -> rounds.  After the first round (gradient at
    the initial point) every chain asks for normal variates; a chain that never receives work stays where it is; the
    status words, the per-chain step of the next lane and the pinned buffers behave as include/b200glm.h says."""
    from stan_b200.model import GLMModel
    d = make_glm_data("bernoulli_logit", 2_000, 6)
    m = GLMModel("bernoulli_logit", d["X"], d["y"])
    L, h, P = _capi.lib(), m.h, m.P
    cfg = _capi.NutsConfig(max_depth=6, num_warmup=0, num_samples=3, w_num_warmup=0, w_init_buffer=0, w_term_buffer=0,
                           w_base_window=0, w_size0=0, w_next0=0xFFFFFFFF, max_deltaH=1000.0, delta=0.8, gamma=0.05,
                           kappa=0.75, t0=10.0, stepsize_jitter=0.5)
    assert L.b200glm_nuts_round(h, 1, (C.c_int32 * 1)(0)) == _capi.INVALID        # not reserved yet
    assert L.b200glm_nuts_reserve(h, 3, C.byref(cfg)) == _capi.OK
    dp = C.POINTER(C.c_double)
    normals, unif, draws, metric = dp(), dp(), dp(), dp()
    status = C.POINTER(_capi.NutsStatus)()
    assert L.b200glm_nuts_buffers(h, C.byref(normals), C.byref(unif), C.byref(status), C.byref(draws),
                                  C.byref(metric)) == _capi.OK
    rng = np.random.default_rng(1)
    q0 = rng.normal(0, 0.1, size=(3, P))
    ones = np.ones(P)
    for c in range(3):
        assert L.b200glm_nuts_init_chain(h, c, q0[c].ctypes.data_as(dp), ones.ctypes.data_as(dp), 0.05) == _capi.OK
        assert status[c].phase == 1 and status[c].iter == 0 and status[c].need_normals == 0
    assert L.b200glm_nuts_init_chain(h, 3, q0[0].ctypes.data_as(dp), ones.ctypes.data_as(dp), 0.05) == _capi.INVALID
    lanes = (C.c_int32 * 2)(0, 2)                                                   # chain 1 is left out
    assert L.b200glm_nuts_round(h, 2, lanes) == _capi.OK
    assert [status[c].phase for c in range(3)] == [2, 1, 2]                         # init_stepsize next; chain 1 untouched
    assert status[0].need_normals == 1 and status[2].need_normals == 1
    # drive chains 0 and 2 to the end with numpy randomness (any stream is a valid sampler; the reference's stream is
    # the driver's business): 3 transitions each
    n_rounds = 0
    STRIDE = 72                                 # a chain's row of `uniforms`: the 64-entry ring, then the jitter variate
    u_jitter = {0: None, 2: None}
    while any(status[c].phase not in (5, 6) for c in (0, 2)) and n_rounds < 2000:
        for c in (0, 2):
            u = rng.random(STRIDE)
            if status[c].need_normals:
                z = rng.standard_normal(P)
                C.memmove(C.addressof(normals.contents) + 8 * c * P, z.ctypes.data, 8 * P)
                if status[c].phase == 4:
                    u_jitter[c] = u[64]         # read when the transition starts
            C.memmove(C.addressof(unif.contents) + 8 * c * STRIDE, u.ctypes.data, 8 * 64)
            if status[c].need_normals:
                C.memmove(C.addressof(unif.contents) + 8 * (c * STRIDE + 64), u[64:].ctypes.data, 8)
        live = [c for c in (0, 2) if status[c].phase not in (5, 6)]
        assert L.b200glm_nuts_round(h, len(live), (C.c_int32 * len(live))(*live)) == _capi.OK
        n_rounds += 1
    assert [status[c].phase for c in range(3)] == [5, 1, 5]
    assert status[0].iter == 3 and status[2].iter == 3 and status[0].adapt_done == 1
    assert status[0].eps_nom == 1.0            # num_warmup == 0: complete_adaptation leaves exp(0) (reference quirk)
    row = np.ctypeslib.as_array(draws, shape=(3, 3 * P + 8))
    assert np.all(np.isfinite(row[[0, 2]])) and row[0, P + 7] == 2 and row[0, P + 4] >= 1
    for c in (0, 2):                           # stepsize__ = eps_nom (1 + jitter (2 u - 1)), base_hmc.hpp:195-200
        assert row[c, P + 2] == 1.0 * (1.0 + 0.5 * (2.0 * u_jitter[c] - 1.0))
    lp, g = m.log_prob_grad(row[0, :P])
    assert abs(lp - row[0, P]) < 1e-9 * abs(lp)                                     # lp__ of the draw is the model's
    assert np.max(np.abs(row[0, 2 * P + 8:] + g)) < 1e-9 * np.max(np.abs(g))        # ... and its gradient columns (of V = -lp)
    energy = -lp + 0.5 * np.sum(row[0, P + 8:2 * P + 8] ** 2)                       # H = V + p.p / 2 (unit metric)
    assert abs(energy - row[0, P + 6]) < 1e-9 * abs(energy)
    m.close()
