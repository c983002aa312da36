#!/usr/bin/env python
"""How long does ONE batched leapfrog take as a function of the number of live lanes?

The lock-step NUTS driver (stan_b200/cpp/b200/batched_nuts.hpp) serves whatever chains are parked, so late in
a run the batches are small; this probe times b200glm_leapfrog_batched (host arrays in and out, as the driver
calls it) for n = 4 ... 1024 lanes of a 1024-slot handle at BASELINE configs[2]'s shape.

    python tools/batch_tail_probe.py [--rows 1000000] [--cols 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--cols", type=int, default=200)
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--lanes", type=int, nargs="*", default=[4, 8, 16, 17, 32, 48, 64, 65, 128, 256, 512, 1024])
    args = ap.parse_args()
    import torch
    from stan_b200 import GLMModel
    dev = torch.device("cuda", 0)
    N, K, C = args.rows, args.cols, args.chains
    g = torch.Generator(device=dev).manual_seed(20261017)
    X = torch.randn((K, N), generator=g, device=dev, dtype=torch.float64)
    beta = torch.randn(K, generator=g, device=dev, dtype=torch.float64) / K ** 0.5
    y = 0.3 + beta @ X + torch.randn(N, generator=g, device=dev, dtype=torch.float64)
    m = GLMModel("normal_id", X.data_ptr(), y.data_ptr(), data_on_device=True, N=N, K=K, ldx=N, device=0)
    del X, y
    m.batch_reserve(C)
    P = m.num_params_r()
    rng = np.random.default_rng(11)
    q0, p0 = 0.05 * rng.standard_normal((C, P)), rng.standard_normal((C, P))
    lp0, g0, st = m.log_prob_grad_batched(q0)
    m.set_state_batched(q0, p0, -g0, -lp0)
    out = []
    for n in args.lanes:
        lanes = rng.permutation(C)[:n].astype(np.int32)
        eps = np.full(n, 1e-5)
        for _ in range(3):
            m.leapfrog_batched(eps, lanes)
        reps = 20 if n <= 128 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            m.leapfrog_batched(eps, lanes)
        ms = (time.perf_counter() - t0) / reps * 1e3
        out.append({"lanes": n, "ms_per_batch": ms, "lane_evals_per_s": n / ms * 1e3})
        print(f"lanes {n:5d}: {ms:8.3f} ms per batched leapfrog  ({n / ms * 1e3:9.0f} gradient evaluations/s)", flush=True)
    print(json.dumps({"workload": f"normal_id N={N} K={K}, {C} chain slots", "batches": out}))
    m.close()


if __name__ == "__main__":
    main()
