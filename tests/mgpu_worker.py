"""Worker for tests/test_multigpu_gpu.py: one rank per GPU, row-sharded model; the likelihood partials
are combined either inside the gradient launch through peer mailboxes (MGPU_TRANSPORT=peer) or by one
NCCL all-reduce per gradient (MGPU_TRANSPORT=nccl).
Launched by torch.distributed.run; prints 'MGPU-OK' from rank 0 when every check passes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stan_b200 import GLMModel, make_glm_data  # noqa: E402
from stan_b200.synth import shard_rows  # noqa: E402


def nuts_through_the_reference_service(rank, world, local, dev):
    """The reference's unmodified hmc_nuts_diag_e_adapt driving a ROW-SHARDED b200::glm_model: every rank
    runs the same deterministic host code on the same seed, every gradient / fused leapfrog is one launch
    per rank with the exchange inside it.  Draws must be bitwise identical on all ranks and agree with the
    unsharded model (same seed) until rounding is amplified."""
    from stan_b200 import stan_service
    if not stan_service.available():
        return []
    failures = []
    N, K = 60_001, 12
    d = make_glm_data("bernoulli_logit", N, K)
    r0, r1 = shard_rows(N, rank, world)
    kw = dict(num_chains=1, seed=77, num_warmup=120, num_samples=80, delta=0.8, num_threads=1)
    m = stan_service.StanGLM("bernoulli_logit", d["X"][r0:r1], d["y"][r0:r1], device=local, n_slots=1,
                             rank=rank, world=world, N_total=N)
    m.connect_peers_torch(dist, dev)
    res = m.nuts(**kw)
    c = m.counters()
    m.close()
    buf = torch.tensor(np.concatenate([res["draws"].ravel(), res["warmup_draws"].ravel()]), device=dev)
    ref = buf.clone()
    dist.broadcast(ref, 0)
    if not torch.equal(buf, ref):
        failures.append("sharded NUTS: ranks hold different draws")
    if rank == 0:
        one = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"], device=local, n_slots=1)
        full = one.nuts(**kw)
        one.close()
        a, b = res["warmup_draws"][:, :5, :], full["warmup_draws"][:, :5, :]
        if not (np.array_equal(a[:, :, 3:6], b[:, :, 3:6]) and np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6):
            failures.append("sharded NUTS: first draws differ from the unsharded model")
        ma, mb = res["draws"][:, :, 7:].mean(axis=(0, 1)), full["draws"][:, :, 7:].mean(axis=(0, 1))
        sd = full["draws"][:, :, 7:].std(axis=(0, 1))
        if not np.all(np.abs(ma - mb) < 1.0 * sd):       # 80 draws each: a loose sanity bound, not the MCSE bar
            failures.append("sharded NUTS: posterior means off")
        if c["leapfrogs"] < res["draws"][:, :, 4].sum():
            failures.append("sharded NUTS: leapfrogs did not run on the device")
        print(f"sharded NUTS world={world}: {int(c['leapfrogs'])} fused leapfrog launches per rank, "
              f"max |mean diff|/sd = {float(np.max(np.abs(ma - mb) / sd)):.2f}", flush=True)
    return failures


def class_outcome_models_sharded(rank, world, local, dev):
    """ordered_logistic / categorical_logit with the rows sharded over the ranks (exchange inside glm_class_kernel's
    launch): == the unsharded oracle, replicated results bitwise identical on every rank, leapfrog included."""
    failures = []
    if os.environ.get("MGPU_TRANSPORT", "peer") != "peer":
        return failures                                      # the class kernels exchange through the mailboxes only
    from oracle.oracle import PortOracle
    for fam, N, K, Cn in [("ordered_logistic", 50_001, 12, 5), ("categorical_logit", 40_000, 9, 4)]:
        d = make_glm_data(fam, N, K, n_classes=Cn)
        r0, r1 = shard_rows(N, rank, world)
        m = GLMModel(fam, d["X"][r0:r1], d["y"][r0:r1], device=local, rank=rank, world=world, N_total=N, n_classes=Cn)
        m.connect_peers_torch(dist, dev)
        rng = np.random.default_rng(6)
        th, p0 = 0.2 * rng.standard_normal(m.P), rng.standard_normal(m.P)
        n0 = m.launch_count()
        lp, g = m.log_prob_grad(th)
        if m.launch_count() - n0 != 1:
            failures.append(f"{fam}: a sharded gradient took {m.launch_count() - n0} launches")
        m.set_state(th, p0, -g, -lp)
        for _ in range(3):
            q1, p1, g1, V1 = m.leapfrog(1e-3)
        buf = torch.tensor(np.concatenate([[lp, V1], g, q1, p1]), device=dev)
        ref = buf.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(buf, ref):
            failures.append(f"{fam}: ranks disagree")
        if rank == 0:
            po = PortOracle(fam, d["X"], d["y"], n_classes=Cn)
            lp_r, g_r = po.log_prob_grad(th)
            sc = np.maximum(np.abs(g_r), np.abs(g_r).max())
            e = max(abs(lp - lp_r) / abs(lp_r), float(np.max(np.abs(g - g_r) / sc)))
            q, p, gg, V = th, p0, -g_r, -lp_r
            for _ in range(3):
                q, p, gg, V = po.leapfrog(1e-3, np.ones(m.P), q, p, gg, V)
            e = max(e, float(np.max(np.abs(q1 - q))), abs(V1 - V) / abs(V))
            if not e < 1e-10:
                failures.append(f"{fam} sharded: err {e}")
            print(f"{fam} N={N} K={K} C={Cn} world={world}: max err {e:.2e}", flush=True)
        m.close()
    return failures


def batched_chains_on_row_shards(rank, world, local, dev):
    """Several chains per pass over a ROW-SHARDED X (few-chain FMA kernel for <= 8 lanes, DMMA kernels above): every
    rank sums its rows for all chains, one NCCL all-reduce of the (K + 2) x chains block per batched evaluation.  ==
    the unsharded oracle chain by chain, bitwise identical on every rank, batched leapfrog included; then the
    device-side NUTS driver on row shards == the same driver on the unsharded model (same seeds)."""
    failures = []
    from oracle.oracle import PortOracle
    for fam, N, K, Cn in [("bernoulli_logit", 120_003, 40, 4), ("poisson_log", 60_001, 20, 7), ("normal_id", 90_001, 33, 20)]:
        d = make_glm_data(fam, N, K)
        r0, r1 = shard_rows(N, rank, world)
        m = GLMModel(fam, d["X"][r0:r1], d["y"][r0:r1], device=local, rank=rank, world=world, N_total=N)
        try:
            m.batch_reserve(Cn)
            failures.append(f"{fam}: batch_reserve before comm_init was accepted on a sharded handle")
        except Exception:
            pass
        m.comm_init_torch(dist, dev)
        m.batch_reserve(Cn)
        rng = np.random.default_rng(8)
        th = 0.1 * rng.standard_normal((Cn, m.P))
        p0 = rng.standard_normal((Cn, m.P))
        lp, g, st = m.log_prob_grad_batched(th)
        m.set_state_batched(th, p0, -g, -lp)
        for _ in range(3):
            q1, p1, g1, V1, st1 = m.leapfrog_batched(np.full(Cn, 1e-3))
        buf = torch.tensor(np.concatenate([lp, V1, g.ravel(), q1.ravel(), p1.ravel()]), device=dev)
        ref = buf.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(buf, ref):
            failures.append(f"{fam} batched on shards: ranks disagree")
        if st.any() or np.any(st1):
            failures.append(f"{fam} batched on shards: status {st} {st1}")
        if rank == 0:
            po = PortOracle(fam, d["X"], d["y"])
            e = 0.0
            for c in range(Cn):
                lp_r, g_r = po.log_prob_grad(th[c])
                sc = np.maximum(np.abs(g_r), np.abs(g_r).max())
                e = max(e, abs(lp[c] - lp_r) / abs(lp_r), float(np.max(np.abs(g[c] - g_r) / sc)))
                q, p, gg, V = th[c], p0[c], -g_r, -lp_r
                for _ in range(3):
                    q, p, gg, V = po.leapfrog(1e-3, np.ones(m.P), q, p, gg, V)
                e = max(e, float(np.max(np.abs(q1[c] - q))), abs(V1[c] - V) / abs(V))
            if not e < 1e-10:
                failures.append(f"{fam} batched on shards: err {e}")
            print(f"{fam} N={N} K={K} chains={Cn} batched on row shards, world={world}: max err {e:.2e}", flush=True)
        m.close()
    from stan_b200 import stan_service
    if stan_service.available():
        N, K = 80_001, 10
        d = make_glm_data("bernoulli_logit", N, K)
        r0, r1 = shard_rows(N, rank, world)
        kw = dict(num_chains=4, seed=31, num_warmup=80, num_samples=40, delta=0.8)
        m = stan_service.StanGLM("bernoulli_logit", d["X"][r0:r1], d["y"][r0:r1], device=local, n_slots=1, rank=rank,
                                 world=world, N_total=N)
        m.connect_peers_torch(dist, dev)       # util::initialize's single-chain gradients: in-kernel exchange
        m.comm_init_torch(dist, dev)           # the batched rounds: one all-reduce each
        res = m.nuts_device(**kw)
        m.close()
        buf = torch.tensor(np.concatenate([res["draws"].ravel(), res["warmup_draws"].ravel()]), device=dev)
        ref = buf.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(buf, ref):
            failures.append("device NUTS on row shards: ranks hold different draws")
        if rank == 0:
            one = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"], device=local, n_slots=1)
            full = one.nuts_device(**kw)
            one.close()
            a, b = res["warmup_draws"][:, :5, :], full["warmup_draws"][:, :5, :]
            if not (np.array_equal(a[:, :, 3:6], b[:, :, 3:6]) and np.max(np.abs(a[:, :, 7:] - b[:, :, 7:])) < 1e-6):
                failures.append("device NUTS on row shards: first draws differ from the unsharded model")
            ma, mb = res["draws"][:, :, 7:].mean(axis=(0, 1)), full["draws"][:, :, 7:].mean(axis=(0, 1))
            sd = full["draws"][:, :, 7:].std(axis=(0, 1))
            if not np.all(np.abs(ma - mb) < 1.0 * sd):
                failures.append("device NUTS on row shards: posterior means off")
            print(f"device NUTS on row shards world={world}: {res['rounds']} rounds, wall {res['wall']:.2f} s "
                  f"(unsharded {full['wall']:.2f} s)", flush=True)
    return failures


def bad_y_is_reported_by_every_rank(rank, world, local, dev):
    """An out-of-range y held by ONE shard (the reference checks y on every call, poisson_log_glm_lpmf.hpp:84) must
    make every rank return the domain error, not only the rank that holds it -- otherwise the replicated host code
    diverges and the next exchange times out."""
    from stan_b200.model import DomainError
    failures = []
    N, K = 40_000, 8
    d = make_glm_data("poisson_log", N, K)
    y = d["y"].copy()
    y[N - 3] = -1                                            # lives in the last rank's shard only
    r0, r1 = shard_rows(N, rank, world)
    m = GLMModel("poisson_log", d["X"][r0:r1], y[r0:r1], device=local, rank=rank, world=world, N_total=N)
    if os.environ.get("MGPU_TRANSPORT", "peer") == "peer":
        m.connect_peers_torch(dist, dev)
    else:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(GLMModel.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        m.comm_init(uid.cpu().numpy().tobytes())
    th = np.zeros(m.num_params_r())
    for _ in range(2):                                       # twice: the ranks must stay in step after the error
        try:
            m.log_prob_grad(th)
            failures.append(f"rank {rank}: out-of-range y in another shard was not reported")
        except DomainError:
            pass
    m.close()
    return failures


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = []
    for fam, N, K, G in [("bernoulli_logit", 200_003, 100, 0), ("poisson_log", 150_000, 50, 1000),
                         ("normal_id", 90_001, 33, 0), ("binomial_logit", 70_001, 21, 0),
                         ("neg_binomial_2_log", 80_003, 17, 0), ("neg_binomial_2_log", 40_000, 9, 37)]:
        d = make_glm_data(fam, N, K, G)                     # every rank builds the same global data
        r0, r1 = shard_rows(N, rank, world)
        grp = None if d["group"] is None else d["group"][r0:r1]
        kw = {"trials": d["trials"][r0:r1]} if fam == "binomial_logit" else {}
        m = GLMModel(fam, d["X"][r0:r1], d["y"][r0:r1], grp, G, device=local, rank=rank, world=world, N_total=N, **kw)
        if os.environ.get("MGPU_TRANSPORT", "peer") == "peer":
            m.connect_peers_torch(dist, dev)
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                uid = torch.frombuffer(bytearray(GLMModel.comm_unique_id()), dtype=torch.uint8).to(dev)
            dist.broadcast(uid, 0)
            m.comm_init(uid.cpu().numpy().tobytes())
        launches0 = m.launch_count()
        P = m.num_params_r()
        rng = np.random.default_rng(5)
        th = 0.1 * rng.standard_normal(P)
        lp, g = m.log_prob_grad(th)
        if os.environ.get("MGPU_TRANSPORT", "peer") == "peer" and G == 0 and m.launch_count() - launches0 != 1:
            failures.append(f"{fam}: a sharded gradient took {m.launch_count() - launches0} launches, expected 1")
        lpd = m.log_prob(th, False, True)
        p0 = rng.standard_normal(P)
        m.set_state(th, p0, -g, -lp)
        for _ in range(3):
            q1, p1, g1, V1 = m.leapfrog(1e-3)
        # replicated theta: every rank must hold bitwise identical results
        buf = torch.tensor(np.concatenate([[lp, lpd, V1], g, q1, p1]), device=dev)
        ref = buf.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(buf, ref):
            failures.append(f"{fam}: ranks disagree")
        if rank == 0:
            from oracle.oracle import PortOracle
            po = PortOracle(fam, d["X"], d["y"], d["group"], G,
                            **({"trials": d["trials"]} if fam == "binomial_logit" else {}))
            lp_r, g_r = po.log_prob_grad(th)
            sc = np.maximum(np.abs(g_r), np.abs(g_r).max())
            e = max(abs(lp - lp_r) / abs(lp_r), float(np.max(np.abs(g - g_r) / sc)),
                    abs(lpd - po.log_prob(th, False, True)) / abs(lpd))
            q, p, gg, V = th, p0, -g_r, -lp_r
            for _ in range(3):
                q, p, gg, V = po.leapfrog(1e-3, np.ones(P), q, p, gg, V)
            e = max(e, float(np.max(np.abs(q1 - q))), abs(V1 - V) / abs(V))
            if not e < 1e-10:
                failures.append(f"{fam}: err {e}")
            print(f"{fam} N={N} K={K} G={G} world={world}: max err {e:.2e}", flush=True)
        m.close()
    failures += class_outcome_models_sharded(rank, world, local, dev)
    failures += batched_chains_on_row_shards(rank, world, local, dev)
    failures += bad_y_is_reported_by_every_rank(rank, world, local, dev)
    failures += nuts_through_the_reference_service(rank, world, local, dev)
    n_fail = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(n_fail)                     # a rank other than 0 may be the one that saw a mismatch
    if failures:
        print(f"rank {rank}: " + "; ".join(failures), flush=True)
    if rank == 0:
        print("MGPU-OK" if int(n_fail.item()) == 0 else "MGPU-FAIL " + "; ".join(failures), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
