#!/usr/bin/env python
"""bench.py -- gradient evaluations per second of the GLM log-density + gradient hot path.

    python bench.py --gpus 1 --steps 50 --warmup 5                       # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1       # the reference's CPU path

A "step" is one gradient evaluation (one fused leapfrog launch: half p, full q, log-density +
gradient over all rows, half p).  Workload at every N: BASELINE.json configs[1], logistic
regression bernoulli_logit_glm N=10M K=100 fp64; with N GPUs the 10M rows are sharded by rows
(strong scaling) and the likelihood partials are combined by one NCCL all-reduce per gradient.

JSON keys: see the task contract.  `value` = device-resident loop (theta never leaves the GPU),
`e2e` = the same metric through the C-ABI call a reference-side caller makes
(b200glm_log_prob_grad with HOST theta / lp / grad buffers, copies inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gradient_evals_per_sec"
UNIT = "grad_evals/s"
N_ROWS = 10_000_000
K_COLS = 100
FAMILY = "bernoulli_logit"


def workload_name(N, K, family=FAMILY, G=0, config=2):
    grp = f" with {G} group intercepts" if G else ""
    return f"{family}_glm N={N} K={K}{grp} fp64, single chain (BASELINE configs[{config - 1}])"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from stan_b200 import GLMModel
    from stan_b200.synth import make_shard_ex

    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N_total, K, G, family = args.rows, args.cols, args.groups, args.family
    if args.weak:
        N_total *= world                      # --weak: --rows is per GPU
    rows, weights = None, None
    if args.balance and world > 1:
        # a row-sharded step waits for its slowest shard: split the rows in proportion to each GPU's measured copy
        # bandwidth instead of equally (opt-in)
        from stan_b200.synth import shard_rows_weighted
        a = torch.empty(1 << 27, device=dev, dtype=torch.float64)        # 1 GiB
        b = torch.empty_like(a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 0.0
        for i in range(8):
            e0.record()
            b.copy_(a)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                best = max(best, 2 * a.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del a, b
        torch.cuda.empty_cache()
        bw = torch.tensor([best], device=dev, dtype=torch.float64)
        allbw = [torch.empty_like(bw) for _ in range(world)]
        dist.all_gather(allbw, bw)
        weights = [float(t.item()) for t in allbw]
        rows = shard_rows_weighted(N_total, weights, rank)
    X, y, grp, trials, r0, r1 = make_shard_ex(torch, dev, family, N_total, K, G, rank, world, rows=rows)
    n_local = r1 - r0
    torch.cuda.synchronize()
    m = GLMModel(family, X.data_ptr(), y.data_ptr(), grp.data_ptr() if G else None, G, data_on_device=True,
                 N=n_local, K=K, ldx=n_local, device=local_rank, rank=rank, world=world, N_total=N_total,
                 trials=trials.data_ptr() if trials is not None else None)
    if world > 1 and args.collective == "peer":
        m.connect_peers_torch(dist, dev)      # in-kernel exchange through peer mailboxes (NVLink), no NCCL call
    elif world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(GLMModel.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        m.comm_init(bytes(uid.cpu().numpy().tobytes()))

    # CPU-baseline sample is taken before X is released (rank 0, N=1 only)
    sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(args.cpu_sample_rows, n_local)
        sample = (X[:, :ns].cpu().numpy().T, y[:ns].cpu().numpy(), grp[:ns].cpu().numpy() if G else None,
                  trials[:ns].cpu().numpy() if trials is not None else None)
    del X, y, grp, trials
    torch.cuda.empty_cache()

    P = m.num_params_r()
    rng = np.random.default_rng(11)
    q0 = 0.05 * rng.standard_normal(P)
    p0 = rng.standard_normal(P)
    lp0, g0 = m.log_prob_grad(q0)
    m.set_state(q0, p0, -g0, -lp0)
    eps = 1e-4
    stream = torch.cuda.ExternalStream(m.stream_ptr(0), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop: `value` ----
    for _ in range(args.warmup):
        m.leapfrog_async(eps)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    launches0 = m.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        m.leapfrog_async(eps)
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = m.launch_count() - launches0
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- end to end through the host-facing C-ABI call: `e2e` ----
    th = q0.copy()
    for _ in range(max(3, args.warmup)):
        m.log_prob_grad(th)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        th[0] = q0[0] + 1e-6 * i
        lp, g = m.log_prob_grad(th)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_ms = (t1 - t0) * 1000.0
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = args.steps / (e2e_ms / 1000.0)

    if rank != 0:
        m.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, which = measured_peaks()
    bytes_per_launch = m.bytes_per_gradient()           # this rank's shard: 8*N*K + 4*N
    achieved = bytes_per_launch / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": which,
                "kernel": f"glm_wide_kernel<{family}>" if K > 256 else f"glm_fused_kernel<{family}>", "algorithmic_bytes_per_launch": bytes_per_launch,
                "avg_launch_ms": ms_per_step}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            for ent in tj.get("entries", [tj]):
                if ent.get("N") == n_local and ent.get("K") == K and ent.get("family", FAMILY) == family:
                    roofline["traffic"] = ent.get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu_baseline = None
    if sample is not None:
        cpu_baseline = cpu_baseline_leg(sample[0], sample[1], N_total, threads=1, evals=args.cpu_evals,
                                        family=family, group=sample[2], G=G, trials=sample[3])
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N_total, K, family, G, args.config), "rows_total": N_total,
                   "rows_per_gpu": n_local, "cols": K, "groups": G,
                   "sharding": f"rows x{world}" + (f" weighted by per-GPU copy bandwidth {[round(w) for w in weights]} GB/s"
                                                   if weights else "") + ((", likelihood partials exchanged inside the gradient launch (peer mailboxes over NVLink)"
                                                    if args.collective == "peer" else
                                                    ", one NCCL all-reduce of P+2 doubles per gradient") if world > 1 else ""),
                   "l2": f"X shard {bytes_per_launch / 1e9:.2f} GB >> 126 MB L2, no flush needed",
                   "step": "one fused leapfrog launch (device-resident theta)"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * P, "d2h_bytes_per_step": 8 * (P + 2),
                "call": "b200glm_log_prob_grad(host theta) -> host lp, grad"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(out))
    m.close()
    if world > 1:
        dist.destroy_process_group()


def run_b200_batched(args):
    """BASELINE configs[2]: normal_id_glm N=1M K=200, 1024 batched chains on one B200.  A step = one
    batched leapfrog (every chain advances once = `chains` gradient evaluations): fp64 DMMA GEMM pair."""
    import torch
    from stan_b200 import GLMModel
    rank, local_rank, world = dist_env()
    if world != 1:
        raise SystemExit("--config 3 is a single-GPU configuration")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N, K, C = args.rows, args.cols, args.chains
    g = torch.Generator(device=dev).manual_seed(20261017)
    X = torch.randn((K, N), generator=g, device=dev, dtype=torch.float64)
    beta = torch.randn(K, generator=g, device=dev, dtype=torch.float64) / K ** 0.5
    y = 0.3 + beta @ X + torch.randn(N, generator=g, device=dev, dtype=torch.float64)
    m = GLMModel("normal_id", X.data_ptr(), y.data_ptr(), data_on_device=True, N=N, K=K, ldx=N, device=local_rank)
    sample = None
    if not args.no_cpu_baseline:
        ns = min(args.cpu_sample_rows, N)
        sample = (X[:, :ns].cpu().numpy().T, y[:ns].cpu().numpy())
    del X, y
    torch.cuda.empty_cache()
    m.batch_reserve(C)
    P = m.num_params_r()
    rng = np.random.default_rng(11)
    q0 = 0.05 * rng.standard_normal((C, P))
    p0 = rng.standard_normal((C, P))
    lp0, g0, st = m.log_prob_grad_batched(q0)
    assert not st.any()
    m.set_state_batched(q0, p0, -g0, -lp0)
    eps = 1e-5
    stream = torch.cuda.ExternalStream(m.batch_stream_ptr(), device=dev)
    for _ in range(args.warmup):
        m.leapfrog_batched_async(C, eps)
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)
    launches0 = m.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        m.leapfrog_batched_async(C, eps)
    e1.record(stream)
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_per_step = e0.elapsed_time(e1) / args.steps
    launches = m.launch_count() - launches0
    clk = clocks.stop(t_wall0, t_wall1)
    value = C * 1000.0 / ms_per_step
    # e2e: host thetas in, host lp/grad out, every step
    th = q0.copy()
    for _ in range(3):
        m.log_prob_grad_batched(th)
    torch.cuda.synchronize()
    ne = max(3, args.steps // 4)
    t0 = time.perf_counter()
    for i in range(ne):
        th[0, 0] = q0[0, 0] + 1e-6 * i
        m.log_prob_grad_batched(th)
    e2e_value = C * ne / (time.perf_counter() - t0)
    flops = 4.0 * N * K * C
    peak = 37.18   # TFLOP/s: register-resident DMMA.8x8x4 loop measured on this pool's B200 (profiles/r1_fp64_peak_microbench.txt)
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    cpu_baseline = None
    if sample is not None:
        cpu_baseline = cpu_baseline_leg(sample[0], sample[1], N, threads=1, evals=args.cpu_evals, family="normal_id")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"normal_id_glm N={N} K={K} fp64, {C} batched chains (BASELINE configs[2])",
                   "rows_total": N, "cols": K, "chains": C,
                   "l2": f"X {8e-9 * N * K:.2f} GB >> 126 MB L2, no flush needed",
                   "step": "one batched leapfrog: begin + fused DMMA GEMM pair + slice reduce + finish (device-resident state)"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * P * C, "d2h_bytes_per_step": 8 * (P + 2) * C,
                "call": "b200glm_log_prob_grad_batched(host thetas) -> host lp, grad"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": "measured fp64 DMMA microbenchmark (tools/fp64_peak.cu); cuBLAS DGEMM 35.4",
                     "kernel": "glm_batched_kernel<normal_id,13>", "algorithmic_flops_per_launch": flops,
                     "avg_launch_ms": ms_per_step},
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(out))
    m.close()


# ------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------
def cpu_baseline_leg(Xs, ys, N_total, threads=1, evals=5, family=FAMILY, group=None, G=0, trials=None):
    from oracle.oracle import PortOracle, RefOracle
    cls = RefOracle if RefOracle.available() else PortOracle
    orc = cls(family, Xs, ys, group, G, **({"trials": trials} if trials is not None else {}))
    ns, K = Xs.shape
    th = 0.05 * np.random.default_rng(11).standard_normal(orc.P)
    orc.log_prob_grad(th)     # warm
    if threads <= 1:
        t0 = time.perf_counter()
        for _ in range(evals):
            orc.log_prob_grad(th)
        dt = (time.perf_counter() - t0) / evals
        rate_sample = 1.0 / dt
    else:
        # the reference has no within-chain threading for GLMs: T chains evaluate concurrently
        def work():
            for _ in range(evals):
                orc.log_prob_grad(th)
        ts = [threading.Thread(target=work) for _ in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = (time.perf_counter() - t0)
        rate_sample = threads * evals / dt
    value = rate_sample * ns / N_total
    out = {"value": value, "unit": UNIT, "cores": threads, "kind": cls.kind,
           "sample": f"first {ns} of {N_total} rows, {evals} evaluations of stan::model::log_prob_grad per thread, "
                     f"rate scaled by rows ({ns}/{N_total})",
           "sample_evals_per_sec": rate_sample}
    if getattr(orc, "isa", None):
        out["isa"] = orc.isa
    return out


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    ns = min(args.cpu_sample_rows, args.rows)
    rng = np.random.Generator(np.random.Philox(key=[20261017, 0]))
    Xs = np.asfortranarray(rng.standard_normal((args.cols, ns)).T)
    beta = np.random.Generator(np.random.Philox(key=[20261017, 1])).standard_normal(args.cols) / np.sqrt(args.cols)
    ys = (rng.random(ns) < 1.0 / (1.0 + np.exp(-(0.3 + Xs @ beta)))).astype(np.int32)
    threads = os.cpu_count() or 1
    from oracle.oracle import PortOracle, RefOracle
    cls = RefOracle if RefOracle.available() else PortOracle
    orc = cls(FAMILY, Xs, ys)
    th = 0.05 * np.random.default_rng(11).standard_normal(orc.P)

    def step():
        ts = [threading.Thread(target=orc.log_prob_grad, args=(th,)) for _ in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    # bounded sample: shrink the row sample until the whole --steps/--warmup run fits in ~2 minutes of wall time
    budget_s = 120.0
    while True:
        t0 = time.perf_counter()
        step()
        t_step = time.perf_counter() - t0
        if t_step * (args.steps + args.warmup) <= budget_s or ns <= 20_000:
            break
        ns = max(20_000, int(ns * min(0.5, budget_s / (t_step * (args.steps + args.warmup)))))
        Xs, ys = np.asfortranarray(Xs[:ns]), ys[:ns]
        orc = cls(FAMILY, Xs, ys)
    for _ in range(max(0, args.warmup - 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    rate_sample = threads * args.steps / dt
    value = rate_sample * ns / args.rows
    ms_per_step = 1000.0 / value
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": cls.kind,
          "sample": f"{ns} of {args.rows} rows; each step = {threads} concurrent chains (threads) each doing one "
                    f"stan::model::log_prob_grad; rate scaled by rows ({ns}/{args.rows})"}
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.rows, args.cols), "rows_total": args.rows, "cols": args.cols},
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE configs index (1-based): 2 = single chain N=10M K=100 (default, the metric's config), "
                         "3 = normal_id N=1M K=200 with 1024 batched chains, 4 = poisson N=50M K=50 with 1000 group "
                         "intercepts (row-sharded), 5 = bernoulli K=1000, 8M rows per GPU (weak scaling; the stated "
                         "N=200M x K=1000 = 1.6 TB does not fit 8 x 180 GB)")
    ap.add_argument("--family", default=None)
    ap.add_argument("--groups", type=int, default=None)
    ap.add_argument("--weak", action="store_true", help="--rows is per GPU (weak scaling)")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: split the rows in proportion to each GPU's measured copy bandwidth (default: equal)")
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the P+2 likelihood partials are summed over ranks")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--cols", type=int, default=None)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--cpu-evals", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    preset = {2: (N_ROWS, K_COLS, FAMILY, 0, False), 3: (1_000_000, 200, "normal_id", 0, False),
              4: (50_000_000, 50, "poisson_log", 1000, False), 5: (8_000_000, 1000, "bernoulli_logit", 0, True)}[args.config]
    args.rows = preset[0] if args.rows is None else args.rows
    args.cols = preset[1] if args.cols is None else args.cols
    args.family = preset[2] if args.family is None else args.family
    args.groups = preset[3] if args.groups is None else args.groups
    args.weak = args.weak or preset[4]
    if args.config == 3 and args.impl == "b200":
        args.steps = args.steps if args.steps is not None else 20
        args.warmup = max(3, args.warmup if args.warmup is not None else 3)
        return run_b200_batched(args)
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 200
        args.warmup = max(3, args.warmup if args.warmup is not None else 10)
        run_b200(args)


if __name__ == "__main__":
    main()
