#!/usr/bin/env python
"""bench_nuts.py -- ESS/s and gradient-evals/s INSIDE NUTS (BASELINE metric, second half).

Runs the reference's unmodified stan::services::sample::hmc_nuts_diag_e_adapt
  (a) on b200::glm_model (GPU, stan_b200/lib/libb200stan.so) and
  (b) on the reference CPU model (oracle/_ref), same seeds, same settings,
and reports, per arm: wall time, gradient evaluations (sum n_leapfrog__ + transitions), grad evals/s,
min/median ESS over parameters (stan::analyze::ess), ESS/s, and the posterior z-scores between arms.

    python bench_nuts.py --config 1                 # N=10k K=20, 4 chains, 1000+1000 (BASELINE configs[0])
    python bench_nuts.py --config 2 --ref-iters 0   # N=10M K=100 on the GPU; CPU arm skipped (hours)
    python bench_nuts.py --config 3 --chains 1024 --warmup 150 --samples 100   # batched chains (DMMA path)
    python bench_nuts.py --config 4 --chains 2 --warmup 150 --samples 150      # poisson + 1000 group intercepts
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_nuts.py --config 2
                                                    # the same single chain with the rows sharded over N GPUs
Writes one JSON line; not part of the driver's bench contract (bench.py is).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def summarize(res, Ref, num_chains):
    d = res["draws"]
    P = d.shape[2] - 7
    ess = [Ref.ess(d[:, :, 7 + k].T) for k in range(P)]
    n_grad_sampling = float(d[:, :, 4].sum() + d.shape[0] * d.shape[1])
    n_grad_warm = float(res["warm_leapfrogs"].sum() + res["warmup_draws"].shape[0] * res["warmup_draws"].shape[1])
    if "rounds" in res:   # device-side driver: no gradient per transition start (the sample's gradient is kept); exact count
        n_grad_warm, n_grad_sampling = float(res["lanes"]), 0.0
    return dict(wall_s=res["wall"], grad_evals=n_grad_warm + n_grad_sampling,
                grad_evals_per_s=(n_grad_warm + n_grad_sampling) / res["wall"],
                ess_min=float(np.min(ess)), ess_median=float(np.median(ess)),
                ess_min_per_s=float(np.min(ess)) / res["wall"],
                mean_treedepth=float(d[:, :, 3].mean()), mean_n_leapfrog=float(d[:, :, 4].mean()),
                divergent=int(d[:, :, 5].sum()), stepsize=[float(s) for s in res["stepsize"]])


def config3(args):
    """BASELINE configs[2]: normal_id_glm N=1M K=200, `--chains` (1024) batched chains on one B200 through
    b200::hmc_nuts_diag_e_adapt_batched.  No CPU arm (1024 chains x N=1M on the host would take days)."""
    from oracle.oracle import RefOracle
    from stan_b200 import make_glm_data, stan_service
    N, K = args.rows or 1_000_000, args.cols or 200
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    t0 = time.time()
    d = make_glm_data("normal_id", N, K)
    t_gen = time.time() - t0
    m = stan_service.StanGLM("normal_id", d["X"], d["y"], n_slots=16, device=local)
    run = m.nuts_device if args.driver == "device" else m.nuts_batched
    if world > 1:
        # N GPUs: the chains shard (X replicated, no data-path collective) -- rank r runs chains [c0, c1) with the chain
        # ids they have in the single-GPU run, so the union of the ranks' draws IS that run's set of chains
        import tempfile
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        c0, c1 = rank * args.chains // world, (rank + 1) * args.chains // world
        dist.barrier()
        res = run(num_chains=c1 - c0, init_chain_id=1 + c0, seed=4711, num_warmup=args.warmup, num_samples=args.samples,
                  delta=0.8)
        wall = torch.tensor([res["wall"]], device=torch.device("cuda", local), dtype=torch.float64)
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        tmp = os.path.join(tempfile.gettempdir(), f"bench_nuts_cfg3_{os.environ.get('MASTER_PORT', '0')}")
        os.makedirs(tmp, exist_ok=True)
        keys = ("draws", "warmup_draws", "stepsize", "inv_metric", "warm_leapfrogs")
        np.savez(os.path.join(tmp, f"r{rank}.npz"), **{k: res[k] for k in keys},
                 counts=np.array([res.get("rounds", res.get("batches", 0)), res["lanes"]]))
        dist.barrier()
        if rank != 0:
            m.close()
            dist.destroy_process_group()
            return
        parts = [np.load(os.path.join(tmp, f"r{r}.npz")) for r in range(world)]
        for k in keys:
            res[k] = np.concatenate([p_[k] for p_ in parts], axis=0)
        res["wall"] = float(wall.item())
        res["lanes"] = int(sum(p_["counts"][1] for p_ in parts))
        res["rounds" if args.driver == "device" else "batches"] = int(max(p_["counts"][0] for p_ in parts))
        res.pop("batch_size_hist", None)
        res["batch_size_hist"] = None
        dist.destroy_process_group()
    else:
        res = run(num_chains=args.chains, seed=4711, num_warmup=args.warmup, num_samples=args.samples, delta=0.8)
    s = summarize(res, RefOracle, args.chains)
    s.pop("stepsize")
    s["stepsize_median"] = float(np.median(res["stepsize"]))
    if args.driver == "device":
        s.update(driver="device-side transition + adaptation (b200::hmc_nuts_diag_e_adapt_device)", rounds=res["rounds"],
                 lanes=res["lanes"], mean_lanes_per_round=res["lanes"] / max(res["rounds"], 1),
                 uniforms_generated=res["uniforms"], normal_vectors_generated=res["normal_vectors"])
    else:
        s.update(driver="host tree building, fibers (b200::hmc_nuts_diag_e_adapt_batched)", batches=res["batches"],
                 lanes=res["lanes"], mean_lanes_per_batch=res["lanes"] / max(res["batches"], 1),
                 batch_size_hist=res["batch_size_hist"])
    per_chain = res["warm_leapfrogs"] + res["draws"][:, :, 4].sum(axis=1)
    pc = lambda a: [float(np.percentile(a, q)) for q in (50, 90, 99, 100)]
    s["per_chain_leapfrogs_p50_p90_p99_max"] = pc(per_chain)
    s["per_chain_stepsize_p0_p1_p50_p100"] = [float(np.percentile(res["stepsize"], q)) for q in (0, 1, 50, 100)]
    s["chains_with_mean_treedepth_ge_8"] = int((res["draws"][:, :, 3].mean(axis=1) >= 8).sum())
    truth = np.concatenate([[d["truth"]["alpha"]], d["truth"]["beta"], [1.0]])
    post_mean = res["draws"][:, :, 7:].mean(axis=(0, 1))
    s["max_abs_post_mean_minus_truth"] = float(np.max(np.abs(post_mean - truth)))
    # Known answer: with N >> K and the weak priors of the model the posterior of (alpha, beta) is the least-squares
    # Gaussian N(b_ols, s^2 (A^T A)^-1), A = [1 X], to O(1/N).  Pooled posterior mean / sd of every coefficient against it,
    # and per chain: how many chains have their own mean further than 5 pooled-MCSE-of-one-chain from the pooled mean.
    A = np.column_stack([np.ones(N), d["X"]])
    coef, *_ = np.linalg.lstsq(A, d["y"], rcond=None)
    resid = d["y"] - A @ coef
    s2 = float(resid @ resid) / (N - K - 1)
    se = np.sqrt(s2 * np.diag(np.linalg.inv(A.T @ A)))
    draws = res["draws"][:, :, 7:7 + K + 1]
    pooled_mean, pooled_sd = draws.mean(axis=(0, 1)), draws.reshape(-1, K + 1).std(axis=0)
    s["ols_check"] = {
        "max_abs_z_of_posterior_mean_vs_ols": float(np.max(np.abs(pooled_mean - coef) / se)),
        "posterior_sd_over_ols_se_min_max": [float(np.min(pooled_sd / se)), float(np.max(pooled_sd / se))],
        "sigma_posterior_mean_vs_ols_s": [float(res["draws"][:, :, 7 + K + 1].mean()), float(np.sqrt(s2))],
    }
    chain_means = draws.mean(axis=1)                                   # (chains, K + 1)
    n_draws = draws.shape[1]
    z_chain = np.abs(chain_means - pooled_mean) / (pooled_sd / np.sqrt(max(n_draws / 4.0, 1.0)))   # ESS >= draws / 4
    s["ols_check"]["chains_with_a_mean_beyond_5_mcse"] = int((z_chain.max(axis=1) > 5.0).sum())
    out = {"workload": f"normal_id_glm N={N} K={K}, NUTS diag_e {args.chains} batched chains {args.warmup}+{args.samples} "
                       + ("via b200::hmc_nuts_diag_e_adapt_device (transition and adaptation on the device)"
                          if args.driver == "device" else
                          "via b200::hmc_nuts_diag_e_adapt_batched (single-chain reference service per chain)"),
           "n_gpus": world, "sharding": "single GPU" if world == 1 else f"chains x{world} (X replicated, no collective)",
           "host_threads": os.cpu_count(), "data_gen_s": t_gen, "b200": dict(s, counters=m.counters())}
    m.close()
    print(json.dumps(out))


def config4(args):
    """BASELINE configs[3] on ONE GPU: poisson_log_glm hierarchical count model N=50M K=50 with 1000 group
    intercepts (P = 1052 parameters), `--chains` chains through the unmodified hmc_nuts_diag_e_adapt; data is
    generated on the device (20 GB of X).  No CPU arm (one gradient of this model takes ~10 s on a host core)."""
    import torch
    from oracle.oracle import RefOracle
    from stan_b200 import stan_service
    from stan_b200.synth import make_shard_ex
    N, K, G = args.rows or 50_000_000, args.cols or 50, 1000
    dev = torch.device("cuda", 0)
    t0 = time.time()
    X, y, grp, _, _, _ = make_shard_ex(torch, dev, "poisson_log", N, K, G, 0, 1)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    m = stan_service.StanGLM("poisson_log", X.data_ptr(), y.data_ptr(), grp.data_ptr(), G, device=0,
                             n_slots=max(2, args.chains), data_on_device=True, N=N, K=K, ldx=N)
    del X, y, grp
    torch.cuda.empty_cache()
    t_upload = time.time() - t0
    res = m.nuts(num_chains=args.chains, seed=4711, num_warmup=args.warmup, num_samples=args.samples, delta=0.8,
                 num_threads=args.chains)
    s = summarize(res, RefOracle, args.chains)
    names = ["mu_a", "sigma_a"]
    post = res["draws"][:, :, 7:].mean(axis=(0, 1))
    s["posterior_mean_mu_a_sigma_a"] = [float(post[0]), float(post[1])]
    print(json.dumps({"workload": f"poisson_log_glm N={N} K={K} with {G} group intercepts (P={m.P}), NUTS diag_e "
                                  f"{args.chains} chains {args.warmup}+{args.samples} via unmodified hmc_nuts_diag_e_adapt",
                      "host_threads": os.cpu_count(), "data_gen_s": t_gen, "upload_relayout_s": t_upload,
                      "b200": dict(s, counters=m.counters())}))
    m.close()


def sharded(args, world, rank, local):
    """BASELINE configs[1] on N GPUs: ONE chain (the config is single-chain) through the unmodified
    hmc_nuts_diag_e_adapt, rows of X sharded over the ranks.  Every rank runs the same host code on the same
    seed; each leapfrog is one fused launch per rank with the likelihood partials exchanged inside it."""
    import torch
    import torch.distributed as dist
    from oracle.oracle import RefOracle
    from stan_b200 import stan_service
    from stan_b200.synth import make_shard
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, K = args.rows or 10_000_000, args.cols or 100
    X, y, _, r0, r1 = make_shard(torch, dev, "bernoulli_logit", N, K, 0, rank, world)
    torch.cuda.synchronize()
    m = stan_service.StanGLM("bernoulli_logit", X.data_ptr(), y.data_ptr(), device=local, n_slots=1, rank=rank,
                             world=world, N_total=N, data_on_device=True, N=r1 - r0, K=K, ldx=r1 - r0)
    del X, y
    torch.cuda.empty_cache()
    m.connect_peers_torch(dist, dev)
    res = m.nuts(num_chains=1, seed=4711, num_warmup=args.warmup, num_samples=args.samples, delta=0.8, num_threads=1)
    wall = torch.tensor([res["wall"]], device=dev, dtype=torch.float64)
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    res["wall"] = float(wall.item())
    if rank == 0:
        s = summarize(res, RefOracle, 1)
        print(json.dumps({"workload": f"bernoulli_logit_glm N={N} K={K}, NUTS diag_e 1 chain {args.warmup}+{args.samples} "
                                      f"via unmodified hmc_nuts_diag_e_adapt, rows sharded over {world} GPUs "
                                      "(exchange inside the leapfrog launch)", "n_gpus": world,
                          "b200": dict(s, counters=m.counters())}))
    m.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--ref-iters", type=int, default=-1, help="CPU arm iterations (warmup=samples); 0 skips, -1 same")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--driver", choices=["service", "batched", "device"], default=None,
                    help="service = the unmodified reference service (default for configs 1, 2, 4); batched = host tree "
                         "building over the batched leapfrog (default for config 3); device = transition + adaptation on "
                         "the device (configs 1 and 3)")
    args = ap.parse_args()
    from oracle.oracle import RefOracle
    from stan_b200 import make_glm_data, stan_service

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    if args.config == 3:
        return config3(args)          # under torchrun: the chains shard over the ranks
    if world > 1:
        return sharded(args, world, rank, local)
    if args.config == 4:
        return config4(args)
    N, K = {1: (10_000, 20), 2: (10_000_000, 100)}[args.config]
    N, K = args.rows or N, args.cols or K
    t0 = time.time()
    d = make_glm_data("bernoulli_logit", N, K)
    t_gen = time.time() - t0
    kw = dict(num_chains=args.chains, seed=4711, num_warmup=args.warmup, num_samples=args.samples, delta=0.8,
              num_threads=args.chains)
    t0 = time.time()
    m = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"], n_slots=max(8, args.chains))
    t_upload = time.time() - t0
    if args.driver in ("device", "batched"):
        kd = {k: v for k, v in kw.items() if k != "num_threads"}
        dev = (m.nuts_device if args.driver == "device" else m.nuts_batched)(**kd)
    else:
        dev = m.nuts(**kw)
    how = {"device": "b200::hmc_nuts_diag_e_adapt_device (transition and adaptation on the device)",
           "batched": "b200::hmc_nuts_diag_e_adapt_batched (host tree building, batched leapfrog)"}.get(
               args.driver, "unmodified hmc_nuts_diag_e_adapt")
    out = {"workload": f"bernoulli_logit_glm N={N} K={K}, NUTS diag_e {args.chains} chains {args.warmup}+{args.samples} "
                       f"via {how}", "host_threads": os.cpu_count(),
           "data_gen_s": t_gen, "upload_relayout_s": t_upload,
           "b200": dict(summarize(dev, RefOracle, args.chains), counters=m.counters())}
    m.close()
    if args.ref_iters != 0:
        kr = dict(kw)
        if args.ref_iters > 0:
            kr.update(num_warmup=args.ref_iters, num_samples=args.ref_iters)
        ro = RefOracle("bernoulli_logit", d["X"], d["y"])
        ref = ro.nuts(**kr)
        out["reference_cpu"] = dict(summarize(ref, RefOracle, args.chains), threads=args.chains, isa=ro.isa,
                                    iters=f'{kr["num_warmup"]}+{kr["num_samples"]}')
        if kr["num_samples"] == kw["num_samples"]:
            zs = []
            for k in range(dev["draws"].shape[2] - 7):
                a, b = dev["draws"][:, :, 7 + k].T, ref["draws"][:, :, 7 + k].T
                zs.append(abs(a.mean() - b.mean()) / np.hypot(RefOracle.mcse_mean(a), RefOracle.mcse_mean(b)))
                zs.append(abs(a.std(ddof=1) - b.std(ddof=1)) / np.hypot(RefOracle.mcse_sd(a), RefOracle.mcse_sd(b)))
            out["posterior_max_z"] = float(max(zs))
        out["speedup_grad_evals_per_s"] = out["b200"]["grad_evals_per_s"] / out["reference_cpu"]["grad_evals_per_s"]
        out["speedup_ess_per_s"] = out["b200"]["ess_min_per_s"] / out["reference_cpu"]["ess_min_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
