"""GPU suite: the reference's OWN services (unmodified hmc_nuts_diag_e_adapt / base_nuts / diag_e_metric /
adaptation / RNG) driving b200::glm_model, against the same services driving the reference CPU model.

Bars (BASELINE.json north_star): lp/gradient 1e-10 relative; posterior means and variances within
Monte-Carlo standard error (stan::analyze::mcse_mean / mcse_sd).
"""
import numpy as np
import pytest

from conftest import rel_err, rel_err_vec, unhex
from stan_b200 import make_glm_data, theta_points
from stan_b200 import stan_service

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not stan_service.available(), reason="libb200stan.so not built")]
TOL = 1e-10


def ref_oracle():
    from oracle.oracle import RefOracle
    if not RefOracle.available():
        pytest.skip("oracle/_ref not present")
    return RefOracle


@pytest.mark.parametrize("name", ["bern_ragged", "pois_groups", "norm_groups", "norm_small"])
def test_model_api_through_reference_functions(golden, name):
    """stan::model::log_prob_grad (tape), stan::model::gradient (specialisation), Model::log_prob<double>."""
    c = golden[name]
    m = stan_service.StanGLM(c["family"], c["X"], c["y"], c["group"], c["G"])
    for e in c["evals"]:
        th = unhex(e["theta"])
        for key, ref in e["lp_grad"].items():
            lp, g = m.log_prob_grad(th, int(key[0]), int(key[1]))
            assert rel_err(lp, float.fromhex(ref["lp"])) < TOL
            assert rel_err_vec(g, unhex(ref["grad"])) < TOL
        lp, g = m.gradient(th)
        assert rel_err(lp, float.fromhex(e["lp_grad"]["11"]["lp"])) < TOL
        assert rel_err_vec(g, unhex(e["lp_grad"]["11"]["grad"])) < TOL
        for key, ref in e["lp_double"].items():
            assert rel_err(m.log_prob(th, int(key[0]), int(key[1])), float.fromhex(ref)) < TOL
    m.close()


def test_specialised_integrator_matches_reference_integrator():
    """expl_leapfrog<diag_e_metric<b200::glm_model>>::evolve (device-resident) == the reference's evolve."""
    Ref = ref_oracle()
    d = make_glm_data("bernoulli_logit", 20_000, 20)
    m = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"])
    ro = Ref("bernoulli_logit", d["X"], d["y"])
    rng = np.random.default_rng(1)
    q0, p0 = 0.1 * rng.standard_normal(m.P), rng.standard_normal(m.P)
    im = np.exp(0.3 * rng.standard_normal(m.P))
    c0 = m.counters()
    q, p, g, V = m.leapfrog(0.02, im, q0, p0, n_steps=25)
    c1 = m.counters()
    qr, pr, gr, Vr = q0, p0, None, 0.0
    qr, pr, gr, Vr = ro.leapfrog(0.02, im, qr, pr, init=True)
    for _ in range(24):
        qr, pr, gr, Vr = ro.leapfrog(0.02, im, qr, pr, gr, Vr)
    assert rel_err_vec(q, qr) < 1e-9 and rel_err_vec(p, pr) < 1e-9 and rel_err_vec(g, gr) < 1e-8
    assert rel_err(V, Vr) < 1e-10
    # 25 fused launches, state uploaded once (device-resident between steps), 1 init gradient
    assert c1["leapfrogs"] - c0["leapfrogs"] == 25
    assert c1["uploads"] - c0["uploads"] == 1
    assert c1["gradients"] - c0["gradients"] == 1
    m.close()


@pytest.fixture(scope="module")
def nuts_pair():
    """BASELINE configs[0]: bernoulli_logit_glm N=10k K=20, NUTS diag_e, 4 chains, 1000+1000."""
    Ref = ref_oracle()
    d = make_glm_data("bernoulli_logit", 10_000, 20)
    m = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"])
    ro = Ref("bernoulli_logit", d["X"], d["y"])
    kw = dict(num_chains=4, seed=4711, num_warmup=1000, num_samples=1000, delta=0.8, num_threads=4)
    dev = m.nuts(**kw)
    dev["counters"] = m.counters()
    ref = ro.nuts(**kw)
    m.close()
    return Ref, dev, ref


def test_nuts_same_seed_same_first_draws(nuts_pair):
    """Host code and RNG streams are identical, lp/grad agree to ~1e-14, so the chains coincide until
    floating-point differences are amplified: the first warm-up draws must match closely."""
    Ref, dev, ref = nuts_pair
    a, b = dev["warmup_draws"][:, :5, :], ref["warmup_draws"][:, :5, :]
    assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])            # treedepth, n_leapfrog, divergent
    assert np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6
    assert np.max(np.abs(a[:, :, 0] - b[:, :, 0]) / np.abs(b[:, :, 0])) < 1e-9   # lp__


def test_nuts_posterior_within_mcse(nuts_pair):
    Ref, dev, ref = nuts_pair
    P = dev["draws"].shape[2] - 7
    zs = []
    for k in range(P):
        a, b = dev["draws"][:, :, 7 + k].T, ref["draws"][:, :, 7 + k].T   # (draws, chains)
        z_mean = abs(a.mean() - b.mean()) / np.hypot(Ref.mcse_mean(a), Ref.mcse_mean(b))
        z_sd = abs(a.std(ddof=1) - b.std(ddof=1)) / np.hypot(Ref.mcse_sd(a), Ref.mcse_sd(b))
        zs += [z_mean, z_sd]
        assert Ref.rhat(a) < 1.02
        assert Ref.ess(a) > 400
    # 42 comparisons: 4 sigma keeps the family-wise false-alarm rate < 0.3 %
    assert max(zs) < 4.0, zs
    assert np.all(dev["draws"][:, :, 5] == 0)                     # no divergences
    # adapted tuning parameters agree to Monte-Carlo noise
    assert np.all(np.abs(dev["stepsize"] / ref["stepsize"] - 1) < 0.35)


def test_nuts_leapfrogs_ran_on_device(nuts_pair):
    Ref, dev, ref = nuts_pair
    n_lf = dev["draws"][:, :, 4].sum() + dev["warm_leapfrogs"].sum()
    c = dev["counters"]
    # every leapfrog of every transition is one fused device launch (plus init_stepsize's)
    assert c["leapfrogs"] >= n_lf
    assert c["uploads"] < c["leapfrogs"]          # state stayed resident for the rest


def test_chains_do_not_depend_on_thread_to_slot_assignment():
    """Regression: thread-local slot / resident-state caches were keyed by the model's ADDRESS, so a model
    allocated where a destroyed one had lived inherited stale slot bindings and two concurrently running
    chains could share one device-resident leapfrog state.  Draws of a chain are a function of (seed, chain
    id) only: 4 chains on 4 threads must be bitwise identical whether every thread has its own slot, all
    threads share ONE slot, or other models lived (and died) on these threads before."""
    d = make_glm_data("bernoulli_logit", 4_000, 6)
    kw = dict(num_chains=4, seed=31, num_warmup=150, num_samples=100, delta=0.8, num_threads=4)
    runs = []
    for n_slots in (8, 1, 2, 8):
        m = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"], n_slots=n_slots)
        th = np.zeros(m.P)
        m.gradient(th)                      # touches the model on the calling thread before the service does
        runs.append(m.nuts(**kw)["draws"])
        m.close()
    for r in runs[1:]:
        assert np.array_equal(runs[0], r)


# ---------------------------------------------------------------------------------------------------
# batched driver: every chain runs the reference's single-chain service on its own host thread, the
# leapfrog steps of all chains are served by one batched DMMA launch (b200/batched_nuts.hpp)
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def batched_pair():
    Ref = ref_oracle()
    d = make_glm_data("normal_id", 5_000, 12)
    m = stan_service.StanGLM("normal_id", d["X"], d["y"])
    kw = dict(num_chains=16, seed=99, num_warmup=300, num_samples=300, delta=0.8)
    bat = m.nuts_batched(**kw)
    bat["counters"] = m.counters()
    seq = m.nuts(num_threads=4, **kw)
    m.close()
    ro = Ref("normal_id", d["X"], d["y"])
    ref = ro.nuts(num_threads=4, **kw)
    return Ref, bat, seq, ref


def test_batched_driver_matches_unbatched_chains(batched_pair):
    """Same seeds => same RNG streams and the same host code: batched chains coincide with the chains of
    the reference's own multi-chain service (on the same device model) until rounding differences between
    the batched and the single-chain kernel are amplified (the parallel_match property of
    T/unit/services/sample/hmc_nuts_diag_e_adapt_parallel_match_test.cpp, to rounding)."""
    Ref, bat, seq, ref = batched_pair
    a, b = bat["warmup_draws"][:, :5, :], seq["warmup_draws"][:, :5, :]
    assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])
    assert np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6
    assert np.max(np.abs(a[:, :, 0] - b[:, :, 0]) / np.abs(b[:, :, 0])) < 1e-9


def test_batched_driver_posterior_within_mcse(batched_pair):
    Ref, bat, seq, ref = batched_pair
    P = bat["draws"].shape[2] - 7
    zs = []
    for k in range(P):
        a, b = bat["draws"][:, :, 7 + k].T, ref["draws"][:, :, 7 + k].T
        zs.append(abs(a.mean() - b.mean()) / np.hypot(Ref.mcse_mean(a), Ref.mcse_mean(b)))
        zs.append(abs(a.std(ddof=1) - b.std(ddof=1)) / np.hypot(Ref.mcse_sd(a), Ref.mcse_sd(b)))
        assert Ref.rhat(a) < 1.03
    assert max(zs) < 4.0, zs
    assert np.all(bat["draws"][:, :, 5] == 0)


def test_batched_driver_batches_chains(batched_pair):
    Ref, bat, seq, ref = batched_pair
    n_eval = bat["draws"][:, :, 4].sum() + bat["warm_leapfrogs"].sum() + 16 * 600   # leapfrogs + one init gradient each
    assert bat["lanes"] >= n_eval
    # lock-step: far fewer launches than evaluations (16 chains, trees of similar depth)
    assert bat["batches"] < 0.25 * bat["lanes"], (bat["batches"], bat["lanes"])


def test_batched_driver_parallel_match_at_1024_chains():
    """The reference's parallel-match property (T/unit/services/sample/hmc_nuts_diag_e_adapt_parallel_match_test.cpp:
    73-140: a chain of the multi-chain service equals the single-chain service run with that chain's id) at the chain
    count of BASELINE configs[2]: 1024 chains as fibers on a few worker threads, every batch one DMMA launch; chains
    picked across the range are compared with the UNBATCHED service started with the same seed and chain id, and the
    pooled posterior with the data-generating parameters."""
    d = make_glm_data("normal_id", 3_000, 4)
    m = stan_service.StanGLM("normal_id", d["X"], d["y"], n_slots=4)
    kw = dict(seed=2024, num_warmup=60, num_samples=40, delta=0.8)
    bat = m.nuts_batched(num_chains=1024, **kw)
    assert np.all(np.isfinite(bat["draws"])) and bat["draws"].shape[0] == 1024
    assert bat["batches"] < 0.02 * bat["lanes"]                  # ~1024 lanes per launch for most of the run
    for c in (0, 1, 511, 777, 1023):
        one = m.nuts(num_chains=1, init_chain_id=1 + c, num_threads=1, **kw)
        a, b = bat["warmup_draws"][c, :5, :], one["warmup_draws"][0, :5, :]
        assert np.array_equal(a[:, 3:6], b[:, 3:6]), c           # tree depth, leapfrog count, divergence flag
        assert np.max(np.abs(a[:, 7:] - b[:, 7:])) < 1e-6, c     # draws agree until rounding is amplified
    m.close()
    pooled = bat["draws"][:, :, 7:].reshape(-1, bat["draws"].shape[2] - 7)
    truth = np.concatenate([[d["truth"]["alpha"]], d["truth"]["beta"], [1.0]])
    sd = pooled.std(axis=0)
    assert np.all(np.abs(pooled.mean(axis=0) - truth) < 4.0 * sd)   # truth inside the posterior (N = 3000 rows)
    assert bat["draws"][:, :, 5].sum() == 0                       # no divergences


def test_batched_driver_serves_shapes_without_dmma_kernel():
    """Group intercepts are outside the DMMA kernel's scope (G == 0 only): the driver still runs the chains
    in lock-step and serves them lane by lane with the single-chain kernel; posterior == reference CPU."""
    Ref = ref_oracle()
    d = make_glm_data("poisson_log", 3_000, 4, 5)
    m = stan_service.StanGLM("poisson_log", d["X"], d["y"], d["group"], 5)
    kw = dict(num_chains=4, seed=7, num_warmup=200, num_samples=200, delta=0.8)
    bat = m.nuts_batched(**kw)
    m.close()
    ref = Ref("poisson_log", d["X"], d["y"], d["group"], 5).nuts(num_threads=4, **kw)
    a, b = bat["warmup_draws"][:, :3, :], ref["warmup_draws"][:, :3, :]
    assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])
    zs = []
    for k in range(bat["draws"].shape[2] - 7):
        x, y = bat["draws"][:, :, 7 + k].T, ref["draws"][:, :, 7 + k].T
        zs.append(abs(x.mean() - y.mean()) / np.hypot(Ref.mcse_mean(x), Ref.mcse_mean(y)))
    assert max(zs) < 4.5, zs


# ---------------------------------------------------------------------------------------------------
# data ingest + the reference's own known-answer posterior
# ---------------------------------------------------------------------------------------------------
def test_json_ingest_reproduces_the_references_posterior_fixture():
    """b200::glm_model(var_context&, glm_config) fed by the reference's stan::json::json_data from a CmdStan-format
    file, sampled by the unmodified hmc_nuts_diag_e_adapt on the GPU, against the posterior means / SDs the reference
    publishes for this data (pathfinder/util.hpp:494-504) at the reference's bars -- and 40x tighter."""
    import os
    from conftest import ROOT
    from test_oracle import check_against_normal_glm_answers, load_normal_glm_fixture
    X, y, exp = load_normal_glm_fixture()
    pri = {k: v for k, v in exp["priors"].items() if k != "source"}
    path = os.path.join(ROOT, "tests", "golden", "normal_glm_data.json")
    m = stan_service.StanGLM.from_json(path, "normal_id", name_y="Y", **pri)
    assert m.P == X.shape[1] + 2 and m.means_x().size == 0
    # the ingested model is the model built from the same arrays
    m2 = stan_service.StanGLM("normal_id", X, y, **pri)
    th = np.concatenate([[-1.0], [-4, -2, 0, 1, 3], [0.0]]) + 0.01
    lp1, g1 = m.gradient(th)
    lp2, g2 = m2.gradient(th)
    assert lp1 == lp2 and np.array_equal(g1, g2)
    m2.close()
    kw = dict(num_chains=4, seed=2026, num_warmup=500, num_samples=500, delta=0.8, num_threads=4)
    res = m.nuts(**kw)
    m.close()
    means = X.mean(axis=0)
    check_against_normal_glm_answers(res["draws"][:, :, 7:], means, exp)
    # brms-style centring as the Stan program INTENDED it (Xc in the likelihood): the intercept moves by
    # dot(means_X, b), everything else is the same posterior
    mc = stan_service.StanGLM.from_json(path, "normal_id", name_y="Y", center_x=True, **pri)
    assert np.allclose(mc.means_x(), means, rtol=0, atol=1e-15)
    rc = mc.nuts(**kw)["draws"][:, :, 7:]
    mc.close()
    K = X.shape[1]
    flat = rc.reshape(-1, rc.shape[2])
    raw_intercept = flat[:, 0] - flat[:, 1:1 + K] @ means
    i_int = exp["names"].index("Intercept")
    assert abs(raw_intercept.mean() - exp["mean"][i_int]) < 0.25 * exp["sd"][i_int]
    for k in range(K):
        assert abs(flat[:, 1 + k].mean() - exp["mean"][2 + k]) < 0.25 * exp["sd"][2 + k]


def test_json_ingest_error_behaviour(tmp_path):
    """A stanc-generated constructor rejects a data block with missing variables or wrong dimensions
    (var_context::validate_dims -> std::runtime_error / invalid_argument); so does this one."""
    import json as js
    from stan_b200.model import CudaError, InvalidArgument
    p = tmp_path / "bad.json"
    p.write_text(js.dumps({"N": 3, "K": 2, "X": [[1, 2], [3, 4]], "y": [0, 1, 1]}))       # X has 2 rows, N = 3
    with pytest.raises((InvalidArgument, CudaError)):
        stan_service.StanGLM.from_json(str(p), "bernoulli_logit")
    p.write_text(js.dumps({"N": 2, "K": 2, "X": [[1, 2], [3, 4]]}))                          # y missing
    with pytest.raises((InvalidArgument, CudaError)):
        stan_service.StanGLM.from_json(str(p), "bernoulli_logit")
    with pytest.raises(InvalidArgument):
        stan_service.StanGLM.from_json(str(tmp_path / "absent.json"), "bernoulli_logit")
    p.write_text(js.dumps({"N": 2, "K": 2, "X": [[1, 2], [3, 4]], "y": [0, 1]}))
    m = stan_service.StanGLM.from_json(str(p), "bernoulli_logit")
    assert m.P == 3
    lp, g = m.gradient(np.zeros(3))
    assert abs(lp - 2 * np.log(0.5)) < 1e-12
    m.close()


def test_csv_and_json_writers_round_trip_through_stan_csv_reader(tmp_path):
    """Output side of the path: the reference's unique_stream_writer / json_writer receive the GPU run's draws
    and adaptation block; stan::io::stan_csv_reader parses the file back to exactly the in-memory draws."""
    import json as js
    d = make_glm_data("poisson_log", 3_000, 4, 3)
    m = stan_service.StanGLM("poisson_log", d["X"], d["y"], d["group"], 3)
    kw = dict(num_chains=2, seed=5, num_warmup=120, num_samples=60, delta=0.8, num_threads=2)
    mem = m.nuts(**kw)
    paths = m.nuts_csv(str(tmp_path / "out"), **kw)
    m.close()
    for c, path in enumerate(paths):
        csv = stan_service.read_stan_csv(path)
        assert csv["header"][:7] == ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__",
                                     "divergent__", "energy__"]
        # stan_csv_reader rewrites "a.2" to "a[2]" (stan_csv_reader.hpp:20-29)
        assert csv["header"][7:] == ["mu_a", "sigma_a", "a[1]", "a[2]", "a[3]", "beta[1]", "beta[2]", "beta[3]", "beta[4]"]
        assert csv["samples"].shape == (60, 7 + m.P)
        assert np.array_equal(csv["samples"], mem["draws"][c])
        # the adaptation block is a comment written at the stream's default 6 significant digits
        # (base_hmc::write_sampler_stepsize / diag_e_point::write_metric through a std::stringstream)
        assert csv["step_size"] == pytest.approx(mem["stepsize"][c], rel=1e-5)
        assert np.allclose(csv["metric"], mem["inv_metric"][c], rtol=1e-5, atol=0)
        with open(str(tmp_path / f"out_metric_{c + 1}.json")) as f:
            mj = js.load(f)
        assert mj["stepsize"] == pytest.approx(mem["stepsize"][c], rel=1e-12)
        assert np.allclose(mj["inv_metric"], mem["inv_metric"][c], rtol=1e-12, atol=0)


def test_log_prob_propto_through_the_reference_function(golden):
    """stan::model::log_prob_propto<jacobian> (log_prob_propto.hpp:32-52, :75-96) -- the call
    base_hamiltonian::update_potential makes -- on b200::glm_model, both signatures, against the same function on
    the reference CPU model (== the value of log_prob_grad<true, jacobian>)."""
    Ref = ref_oracle()
    for name in ("bern_small", "pois_groups", "norm_small"):
        c = golden[name]
        m = stan_service.StanGLM(c["family"], c["X"], c["y"], c["group"], c["G"])
        orc = Ref(c["family"], c["X"], c["y"], c["group"], c["G"])
        for th in theta_points(m.P, n_random=2, scale=0.2):
            for jac in (True, False):
                want = orc.log_prob_grad(th, True, jac)[0]
                assert rel_err(m.log_prob_propto(th, jac, eigen=False), want) < TOL
                assert rel_err(m.log_prob_propto(th, jac, eigen=True), want) < TOL
        m.close()


def _r_dump(path, **vars_):
    """CmdStan's R-dump data format (what stan::io::dump parses)."""
    def num(v):
        return repr(float(v)) if isinstance(v, (float, np.floating)) else str(int(v))
    with open(path, "w") as f:
        for k, v in vars_.items():
            a = np.asarray(v)
            if a.ndim == 0:
                f.write(f"{k} <- {num(a[()])}\n")
            elif a.ndim == 1:
                f.write(f"{k} <- c({', '.join(num(x) for x in a)})\n")
            else:   # column-major, as R stores matrices
                flat = ", ".join(num(x) for x in a.flatten(order="F"))
                f.write(f"{k} <- structure(c({flat}), .Dim = c({', '.join(str(d) for d in a.shape)}))\n")


def test_dump_ingest_matches_array_construction(tmp_path):
    """the stanc-style constructor fed by the reference's R-dump reader stan::io::dump
    (ST/io/dump.hpp) gives the same model as construction from arrays, and validates the data block the same way."""
    from stan_b200.model import InvalidArgument
    d = make_glm_data("poisson_log", 500, 4, 7)
    p = tmp_path / "data.R"
    _r_dump(str(p), N=500, K=4, X=d["X"], y=d["y"], G=7, group=d["group"])
    m = stan_service.StanGLM.from_dump(str(p), "poisson_log")
    ref = stan_service.StanGLM("poisson_log", d["X"], d["y"], d["group"], 7)
    assert m.P == ref.P == 2 + 7 + 4
    for th in theta_points(m.P, n_random=2, scale=0.2):
        a, b = m.log_prob_grad(th), ref.log_prob_grad(th)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
    m.close()
    ref.close()
    bad = tmp_path / "bad.R"
    _r_dump(str(bad), N=500, K=4, X=d["X"][:, :3], y=d["y"])            # X has the wrong shape
    with pytest.raises(InvalidArgument):
        stan_service.StanGLM.from_dump(str(bad), "poisson_log")


@pytest.mark.parametrize("fam,K,C", [("ordered_logistic", 4, 4), ("categorical_logit", 3, 3)])
def test_class_outcome_models_through_the_unmodified_service(fam, K, C):
    """The two class-outcome GLMs driven by the reference's hmc_nuts_diag_e_adapt through b200::glm_model (names,
    dims, ordered_constrain in write_array): model API == reference CPU model, same-seed first draws, posterior means
    within Monte-Carlo standard error, constrained cut-points come out ordered."""
    Ref = ref_oracle()
    d = make_glm_data(fam, 3_000, K, n_classes=C)
    m = stan_service.StanGLM(fam, d["X"], d["y"], n_classes=C)
    ro = Ref(fam, d["X"], d["y"], n_classes=C)
    assert m.P == ro.P
    for th in theta_points(m.P, n_random=2, scale=0.3):
        for propto in (True, False):
            a, b = m.log_prob_grad(th, propto, True), ro.log_prob_grad(th, propto, True)
            assert rel_err(a[0], b[0]) < TOL and rel_err_vec(a[1], b[1]) < TOL
        g, gr = m.gradient(th), ro.gradient(th)
        assert rel_err(g[0], gr[0]) < TOL and rel_err_vec(g[1], gr[1]) < TOL
    kw = dict(num_chains=4, seed=11, num_warmup=300, num_samples=300, delta=0.8, num_threads=4)
    dev = m.nuts(**kw)
    c = m.counters()
    m.close()
    ref = ro.nuts(**kw)
    a, b = dev["warmup_draws"][:, :4, :], ref["warmup_draws"][:, :4, :]
    assert np.array_equal(a[:, :, 3:6], b[:, :, 3:6])
    assert np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6
    assert c["leapfrogs"] >= dev["draws"][:, :, 4].sum()          # the fused device leapfrog served them
    zs = []
    for k in range(dev["draws"].shape[2] - 7):
        x, y = dev["draws"][:, :, 7 + k].T, ref["draws"][:, :, 7 + k].T
        zs.append(abs(x.mean() - y.mean()) / np.hypot(Ref.mcse_mean(x), Ref.mcse_mean(y)))
        assert Ref.rhat(x) < 1.05
    assert max(zs) < 4.5, zs
    assert dev["draws"][:, :, 5].sum() == 0
    if fam == "ordered_logistic":                                  # write_array applies ordered_constrain
        cuts = dev["draws"][:, :, 7 + K:]
        assert np.all(np.diff(cuts, axis=2) > 0)
