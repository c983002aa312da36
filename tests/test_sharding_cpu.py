"""CPU suite (gloo, world_size 2): the host-side logic of the row-sharded N>1 path.

What the GPU path relies on, checked here without a GPU:
  * shard_rows partitions [0, N) exactly, and every sharding regenerates the same global X/y;
  * the likelihood part of (lp, grad) is additive over row shards, so ONE all-reduce (sum) of the
    per-shard likelihood partials followed by adding the priors once reproduces the unsharded result
    -- the arithmetic b200glm does with ncclAllReduce + finish_kernel, done here with the C oracle
    per shard and torch.distributed(gloo).all_reduce.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stan_b200.synth import make_logistic_shard, shard_rows  # noqa: E402


def test_shard_rows_partition():
    for N in (0, 1, 31, 32, 33, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_rows(N, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == N
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            assert sum(b - a for a, b in ranges) == N


def _worker(rank, world, port, N, K, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle.oracle import PortOracle
    X, y, r0, r1 = make_logistic_shard(torch, torch.device("cpu"), N, K, rank, world, block=1000)
    Xn, yn = X.numpy().T, y.numpy()
    # the 128-byte communicator id travels the same way in bench.py (rank 0 -> everyone)
    uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    assert uid[77].item() == 77
    P = K + 1
    th = 0.1 * np.random.default_rng(3).standard_normal(P)
    shard = PortOracle("bernoulli_logit", Xn, yn)
    empty = PortOracle("bernoulli_logit", np.zeros((0, K)), np.zeros(0, np.int32))
    lp_s, g_s = shard.log_prob_grad(th)
    lp_0, g_0 = empty.log_prob_grad(th)               # priors only
    lik = torch.tensor(np.concatenate([g_s - g_0, [lp_s - lp_0]]))   # likelihood partials of this shard
    dist.all_reduce(lik, op=dist.ReduceOp.SUM)        # the one collective per gradient
    total = lik.numpy()
    lp, g = total[-1] + lp_0, total[:-1] + g_0        # priors added once, identically on every rank
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, r1, Xn, yn))
    if rank == 0:
        Xf = np.concatenate([t[2] for t in gathered], axis=0)
        yf = np.concatenate([t[3] for t in gathered])
        out.put((lp, g, Xf, yf, th))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_gradient_world2_gloo():
    N, K, world = 5003, 7, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, K, out)) for r in range(world)]
    for p in procs:
        p.start()
    lp, g, Xf, yf, th = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # same global data as the unsharded generator
    X1, y1, _, _ = make_logistic_shard(torch, torch.device("cpu"), N, K, 0, 1, block=1000)
    assert np.array_equal(X1.numpy().T, Xf) and np.array_equal(y1.numpy(), yf)
    from oracle.oracle import PortOracle
    lp_ref, g_ref = PortOracle("bernoulli_logit", Xf, yf).log_prob_grad(th)
    assert abs(lp - lp_ref) / abs(lp_ref) < 1e-13
    assert np.max(np.abs(g - g_ref)) / np.max(np.abs(g_ref)) < 1e-13


def test_make_shard_is_sharding_invariant_for_every_family():
    """bench.py --config 4/5 data: any row sharding regenerates the same global X, y, group."""
    from stan_b200.synth import make_shard
    dev = torch.device("cpu")
    for fam, G in (("poisson_log", 7), ("normal_id", 0), ("bernoulli_logit", 3)):
        X, y, g, _, _ = make_shard(torch, dev, fam, 2500, 5, G, 0, 1, block=1000)
        parts = [make_shard(torch, dev, fam, 2500, 5, G, r, 3, block=1000) for r in range(3)]
        assert torch.equal(X, torch.cat([p[0] for p in parts], 1))
        assert torch.equal(y, torch.cat([p[1] for p in parts]))
        if G:
            assert torch.equal(g, torch.cat([p[2] for p in parts])) and 1 <= int(g.min()) and int(g.max()) <= G
    # binomial_logit / neg_binomial_2_log (make_shard_ex also returns the population sizes)
    from stan_b200.synth import make_shard_ex
    for fam, G in (("binomial_logit", 0), ("binomial_logit", 4), ("neg_binomial_2_log", 0)):
        X, y, g, t, _, _ = make_shard_ex(torch, dev, fam, 2500, 5, G, 0, 1, block=1000)
        parts = [make_shard_ex(torch, dev, fam, 2500, 5, G, r, 3, block=1000) for r in range(3)]
        assert torch.equal(X, torch.cat([p[0] for p in parts], 1))
        assert torch.equal(y, torch.cat([p[1] for p in parts])) and int(y.min()) >= 0
        if fam == "binomial_logit":
            assert torch.equal(t, torch.cat([p[3] for p in parts])) and bool((y <= t).all())
        else:
            assert t is None


def test_weighted_row_shards():
    """bench.py --balance: rows split in proportion to per-GPU bandwidth; the blocks stay contiguous, aligned to
    the panel height, and regenerate the same global data as the equal split."""
    from stan_b200.synth import make_shard_ex, shard_rows_weighted
    N = 100_003
    for w in ([1.0], [1, 1, 1], [6.5, 6.1, 6.5, 6.4, 6.5, 6.5, 6.2, 6.5], [1e-3, 1.0], [5, 1, 1, 1, 1, 1, 1, 1]):
        b = [shard_rows_weighted(N, w, r) for r in range(len(w))]
        assert b[0][0] == 0 and b[-1][1] == N
        assert all(b[i][1] == b[i + 1][0] for i in range(len(w) - 1))
        assert all(r0 % 32 == 0 and r1 >= r0 for r0, r1 in b)
        share = np.array([r1 - r0 for r0, r1 in b]) / N
        assert np.max(np.abs(share - np.asarray(w, float) / np.sum(w))) < 64 / N + 1e-12
    with pytest.raises(ValueError):
        shard_rows_weighted(N, [1.0, 0.0], 0)
    dev = torch.device("cpu")
    X, y, _, _, _, _ = make_shard_ex(torch, dev, "bernoulli_logit", 2500, 5, 0, 0, 1, block=1000)
    w = [2.0, 1.0, 3.0]
    parts = [make_shard_ex(torch, dev, "bernoulli_logit", 2500, 5, 0, r, 3, block=1000,
                           rows=shard_rows_weighted(2500, w, r)) for r in range(3)]
    assert torch.equal(X, torch.cat([p[0] for p in parts], 1)) and torch.equal(y, torch.cat([p[1] for p in parts]))
    assert [p[5] - p[4] for p in parts] == [r1 - r0 for r0, r1 in (shard_rows_weighted(2500, w, r) for r in range(3))]
