"""Device-side NUTS, checked WITHOUT a GPU: the product's driver (stan_b200/cpp/b200/device_nuts.hpp) and
per-chain state machine (stan_b200/csrc/nuts_tree.cuh -- the code the kernels run with one warp per chain) are built for
the host by the checker (oracle/ref/nuts_host_backend.hpp: one-lane policy, leapfrog steps by the reference's own
integrator on the reference's own model) and run against stan::services::sample::hmc_nuts_diag_e_adapt on the same seeds.

What is being compared is exactly what moved to the device: the iterative build_tree (base_nuts.hpp:247-352), transition()
(:78-204), init_stepsize (base_hmc.hpp:78-143), dual averaging (stepsize_adaptation.hpp:55-71), the metric windows
(var_adaptation.hpp:17-46, windowed_adaptation.hpp) and the host's engine bookkeeping (normal / uniform variates in the
reference's order, rewound to what the device consumed).  Every draw column is compared: lp__, accept_stat__, stepsize__,
treedepth__, n_leapfrog__, divergent__, energy__, parameters.  Dot products are summed in a different order than Eigen's,
so agreement is to rounding, amplified by the dynamics as the run goes on; the bars below say how far."""
import numpy as np
import pytest

from oracle.oracle import RefOracle
from stan_b200.synth import make_glm_data

pytestmark = pytest.mark.skipif(not RefOracle.available(), reason="oracle/_ref not built (needs /root/reference)")


def _both(fam, N, K, G=0, **kw):
    d = make_glm_data(fam, N, K, G)
    ro = RefOracle(fam, d["X"], d["y"], d["group"], G)
    a, b = ro.nuts(**kw), ro.nuts_device_host(**kw)
    A = np.concatenate([a["warmup_draws"], a["draws"]], axis=1)
    B = np.concatenate([b["warmup_draws"], b["draws"]], axis=1)
    return a, b, A, B


def _err(A, B):
    return np.abs(A - B) / (1.0 + np.abs(A))


@pytest.mark.parametrize("fam,N,K", [("bernoulli_logit", 500, 4), ("normal_id", 400, 3), ("poisson_log", 300, 6)])
def test_whole_run_matches_the_reference_sampler(fam, N, K):
    """150 warm-up (three metric windows, four init_stepsize searches) + 50 sampling iterations, 3 chains: every column of
    every draw, the adapted step size and the adapted metric."""
    kw = dict(num_chains=3, seed=11, init_chain_id=2, num_warmup=150, num_samples=50, stepsize=1.0, max_depth=10, delta=0.8)
    a, b, A, B = _both(fam, N, K, **kw)
    assert np.array_equal(A[:, :, 3:6], B[:, :, 3:6])          # treedepth__, n_leapfrog__, divergent__: same trees
    e = _err(A, B).max(axis=(0, 2))
    assert e[:20].max() < 1e-12 and e.max() < 1e-6            # rounding, amplified over 200 iterations
    assert np.abs(a["stepsize"] - b["stepsize"]).max() < 1e-8
    assert np.abs(a["inv_metric"] - b["inv_metric"]).max() < 1e-8
    # the host generated exactly one vector of normal variates per momentum refresh
    assert b["normal_vectors"] >= 3 * 200


def test_hierarchical_model_matches_until_rounding_is_amplified():
    """a = a[group] with sigma_a: depth-5 trees and chaotic dynamics.  The first iterations agree to the last bit or two,
    the difference then grows smoothly (rounding amplified), never by a jump (which a logic difference would give)."""
    kw = dict(num_chains=2, seed=11, num_warmup=150, num_samples=20, stepsize=1.0, max_depth=10, delta=0.8)
    a, b, A, B = _both("poisson_log", 600, 3, 5, **kw)
    per_iter = _err(A, B).max(axis=(0, 2))
    assert per_iter[:12].max() < 1e-13
    assert per_iter[:30].max() < 1e-8
    first_tree_diff = np.argmax((A[:, :, 4] != B[:, :, 4]).any(axis=0)) if (A[:, :, 4] != B[:, :, 4]).any() else len(per_iter)
    assert first_tree_diff >= 40


@pytest.mark.parametrize("num_warmup,num_samples,max_depth,stepsize", [
    (0, 30, 10, 0.3),      # no warm-up: complete_adaptation still sets the step size to exp(0) = 1 (reference quirk)
    (10, 30, 10, 1.0),     # num_warmup < 20: no metric estimation, dual averaging only
    (60, 20, 10, 1.0),     # windows do not fit: the 15 % / 75 % / 10 % schedule
    (100, 30, 3, 1.0),     # trees cut at max_depth
    (100, 20, 10, 50.0),   # absurd initial step size: init_stepsize halves it many times; early divergences
])
def test_schedules_and_limits(num_warmup, num_samples, max_depth, stepsize):
    kw = dict(num_chains=2, seed=5, num_warmup=num_warmup, num_samples=num_samples, stepsize=stepsize,
              max_depth=max_depth, delta=0.8)
    a, b, A, B = _both("bernoulli_logit", 400, 5, **kw)
    assert np.array_equal(A[:, :, 3:6], B[:, :, 3:6])
    e = _err(A, B).max(axis=(0, 2))
    assert e[:20].max() < 1e-12 and e.max() < 1e-6
    assert np.abs(a["stepsize"] - b["stepsize"]).max() < 1e-8
    assert np.abs(a["inv_metric"] - b["inv_metric"]).max() < 1e-8
    if max_depth == 3:
        assert (A[:, :, 3] == 3).any() and (A[:, :, 4] <= 7).all()
    if num_warmup == 0:
        assert np.all(A[:, :, 2] == 1.0)


def test_stepsize_jitter_is_the_reference_s():
    """stepsize_jitter > 0: sample_stepsize() (base_hmc.hpp:195-200) draws ONE uniform variate per transition, before the
    momentum (base_nuts.hpp:80 precedes :84) -- every later variate of the chain moves by one, so any slip in the host's
    engine bookkeeping shows in the first rows.  stepsize__ is the jittered value; dual averaging acts on the nominal one.
    Values outside (0, 1) are ignored by base_hmc::set_stepsize_jitter."""
    kw = dict(num_chains=2, seed=7, num_warmup=100, num_samples=40, stepsize=1.0, max_depth=10, delta=0.8)
    a, b, A, B = _both("bernoulli_logit", 400, 5, stepsize_jitter=0.4, **kw)
    assert np.array_equal(A[:, :, 3:6], B[:, :, 3:6])
    e = _err(A, B).max(axis=(0, 2))
    assert e[:20].max() < 1e-12 and e.max() < 1e-6
    assert np.abs(a["stepsize"] - b["stepsize"]).max() < 1e-8
    ratio = B[:, 100:, 2] / b["stepsize"][:, None]            # after warm-up: eps / eps_nom in [0.6, 1.4], not constant
    assert ratio.min() >= 0.6 - 1e-9 and ratio.max() <= 1.4 + 1e-9 and ratio.std() > 0.1
    a0, b0, A0, B0 = _both("bernoulli_logit", 400, 5, **kw)
    assert not np.array_equal(A0, A)                          # the jitter changed the reference's chains ...
    a1, b1, A1, B1 = _both("bernoulli_logit", 400, 5, stepsize_jitter=1.5, **kw)
    assert np.array_equal(A1, A0) and np.array_equal(B1, B0)  # ... and an out-of-range value changes nothing


@pytest.mark.parametrize("num_warmup", [10, 120])
def test_initial_inverse_metric_is_used(num_warmup):
    """A non-unit diagonal inverse metric handed in through the init_inv_metric contexts (read_diag_inv_metric,
    hmc_nuts_diag_e_adapt.hpp:76-79): momenta are drawn with it, the sharp momenta of the U-turn checks use it, and with
    num_warmup < 20 it is never replaced -- the chains must stay the reference's, and differ from the unit-metric ones."""
    im = np.exp(np.random.default_rng(2).normal(0, 0.7, 6))          # alpha + 5 betas
    kw = dict(num_chains=2, seed=13, num_warmup=num_warmup, num_samples=30, stepsize=0.5, max_depth=10, delta=0.8)
    a, b, A, B = _both("bernoulli_logit", 400, 5, init_inv_metric=im, **kw)
    assert np.array_equal(A[:, :, 3:6], B[:, :, 3:6])
    e = _err(A, B).max(axis=(0, 2))
    assert e[:20].max() < 1e-12 and e.max() < 1e-6
    assert np.abs(a["inv_metric"] - b["inv_metric"]).max() < 1e-8
    if num_warmup < 20:
        assert np.allclose(b["inv_metric"], im[None, :], rtol=0, atol=1e-15)
    a0, b0, A0, B0 = _both("bernoulli_logit", 400, 5, **kw)
    assert not np.array_equal(B0[:, :5, 7:], B[:, :5, 7:])


@pytest.mark.parametrize("adapt,delta", [
    (dict(gamma=0.1, kappa=0.6, t0=5.0, init_buffer=20, term_buffer=10, window=8), 0.9),    # many short metric windows
    (dict(init_buffer=10, term_buffer=60, window=30), 0.65),     # the last window absorbs the rest (windowed_adaptation.hpp:97-113)
    (dict(gamma=-1.0, kappa=0.0, t0=-3.0), 1.5),                 # out of range: the setters keep 0.05 / 0.75 / 10 / 0.5
    (dict(init_buffer=100, term_buffer=100, window=100), 0.8),   # does not fit 150: the 15 % / 75 % / 10 % schedule
])
def test_adaptation_parameters_of_the_service(adapt, delta):
    """gamma / kappa / t0 / delta (stepsize_adaptation's setters, which ignore out-of-range values) and init_buffer /
    term_buffer / window (windowed_adaptation::set_window_params) other than the defaults: same chains, same adapted step
    size, same adapted metric as the reference service."""
    kw = dict(num_chains=2, seed=17, num_warmup=150, num_samples=30, stepsize=1.0, max_depth=10, delta=delta, adapt=adapt)
    a, b, A, B = _both("poisson_log", 300, 4, **kw)
    assert np.array_equal(A[:, :, 3:6], B[:, :, 3:6])
    e = _err(A, B).max(axis=(0, 2))
    assert e[:20].max() < 1e-12 and e.max() < 1e-6
    assert np.abs(a["stepsize"] - b["stepsize"]).max() < 1e-8
    assert np.abs(a["inv_metric"] - b["inv_metric"]).max() < 1e-8


def test_deep_trees_up_to_max_depth():
    """delta = 0.99 on a 62-parameter hierarchical model: step sizes small enough for trees of depth 9 and 10 (1023
    leapfrogs, ten levels of the explicit subtree stack, up to eleven uniform variates consumed in one round).  Tree depths,
    leapfrog counts and divergence flags of all 210 iterations are the reference's; so is the adapted step size."""
    kw = dict(num_chains=4, seed=23, num_warmup=150, num_samples=60, stepsize=1.0, max_depth=10, delta=0.99)
    a, b, A, B = _both("poisson_log", 200, 2, 60, **kw)
    assert A[:, :, 3].max() == 10 and (A[:, :, 4] == 1023).any()
    assert np.array_equal(A[:, :, 3:6], B[:, :, 3:6])
    assert _err(A, B)[:, :20].max() < 1e-10 and _err(A, B).max() < 1e-5
    assert np.abs(a["stepsize"] - b["stepsize"]).max() < 1e-8


def test_divergent_transitions_are_reproduced():
    """A step size far too large for a sharp posterior, no adaptation to repair it: divergent__ = 1 rows must coincide."""
    kw = dict(num_chains=2, seed=3, num_warmup=0, num_samples=60, stepsize=1.0, max_depth=10, delta=0.8)
    a, b, A, B = _both("poisson_log", 4000, 4, **kw)
    assert A[:, :, 5].sum() > 0
    assert np.array_equal(A[:, :, 3:6], B[:, :, 3:6])
    assert _err(A, B).max() < 1e-6


def _by_tag(lines):
    out = {}
    for line in lines:
        tag, _, rest = line.partition("|")
        out.setdefault(tag, []).append(rest)
    return out


@pytest.mark.parametrize("kw", [
    dict(num_warmup=40, num_samples=20, num_thin=3, save_warmup=False, refresh=10),
    dict(num_warmup=30, num_samples=15, num_thin=1, save_warmup=True, refresh=0),
    dict(num_warmup=0, num_samples=12, num_thin=5, save_warmup=True, refresh=4),
])
def test_writers_and_logger_receive_what_the_reference_service_sends(kw):
    """Everything the sample writer, the diagnostic writer and the logger receive -- column names, thinned rows,
    warm-up rows or not, "Adaptation terminated" / "Step size" / metric lines, the refresh messages, the diagnostic
    writer's q / p / g columns -- line by line against stan::services::sample::hmc_nuts_diag_e_adapt with the same service
    options (timing lines excepted; numbers to 1e-9)."""
    d = make_glm_data("bernoulli_logit", 300, 3)
    ro = RefOracle("bernoulli_logit", d["X"], d["y"])
    A = _by_tag(ro.nuts_transcript(0, num_chains=2, seed=9, **kw))
    B = _by_tag(ro.nuts_transcript(1, num_chains=2, seed=9, **kw))
    assert sorted(A) == sorted(B)
    n_rows = 0
    for tag in A:
        assert len(A[tag]) == len(B[tag]), tag
        if tag == "L":   # the logger is shared by the chains, which run concurrently here and one after the other in
            # the reference arm: the same messages, in another interleaving
            drop = lambda ls: sorted(l for l in ls if "seconds" not in l)
            assert drop(A[tag]) == drop(B[tag])
            continue
        for x, y in zip(A[tag], B[tag]):
            if "seconds" in x or "Adjust your expectations" in x:      # timings
                assert "seconds" in y or "Adjust your expectations" in y
                continue
            if x == y:
                continue
            fx, fy = x.lstrip("#").split(","), y.lstrip("#").split(",")
            if x.startswith("#Step size"):
                fx, fy = [x.split("=")[1]], [y.split("=")[1]]
            assert len(fx) == len(fy), (tag, x, y)
            a, b = np.array(fx, dtype=float), np.array(fy, dtype=float)
            assert np.max(np.abs(a - b) / (1.0 + np.abs(a))) < 1e-9, (tag, x, y)
            n_rows += 1
    # the thinned / unthinned row counts are the reference's (checked above by equal lengths); make sure rows were there
    assert sum(1 for r in A["S0"] if r and r[0] not in "#l") >= (kw["num_samples"] + kw["num_thin"] - 1) // kw["num_thin"]
