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
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md), polled through NVML every few
    milliseconds from a thread (the timed region of a 20-step run is ~20 ms: nvidia-smi's loop mode is too coarse for
    it); falls back to `nvidia-smi -lms` when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines, self.samples = index, None, [], []
        self.nvml, self.stop_flag, self.t = None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    pass
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown")
                else n.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0)),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0)),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap",
                                        getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0))}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons",
                              getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = get_reasons(self.handle) if get_reasons else 0
                self.samples.append((time.time(), sm, mx, [k for k, b in bits.items() if b and (r & b)]))
            except Exception:
                pass
            time.sleep(0.003)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.nvml is not None:
            time.sleep(0.01)
            self.stop_flag = True
            self.t.join(timeout=1.0)
            inside = [s for s in self.samples if t0 is None or t0 <= s[0] <= t1]
            how = "NVML every 3 ms, samples inside the timed region"
            if not inside and self.samples and t0 is not None:   # region shorter than one poll: the nearest samples
                mid = 0.5 * (t0 + t1)
                inside = sorted(self.samples, key=lambda s: abs(s[0] - mid))[:2]
                how = "NVML every 3 ms, the two samples nearest to the timed region"
            if not inside:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML samples"]}
            reasons = sorted({r for s in inside for r in s[3]})
            return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": inside[0][2],
                    "samples": len(inside), "reasons": reasons, "how": how}
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
                "samples": len(sm), "reasons": sorted(reasons), "how": "nvidia-smi -lms 100"}


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
    streamed = bool(args.streamed)
    if streamed:
        m, n_local, N_total, sample, par_sample = build_streamed(torch, dist, dev, args, rank, world, local_rank)
        m_local, par_rows = None, (par_sample[0].shape[0] if par_sample is not None else 0)
        if world > 1 and args.collective == "peer":
            m.connect_peers_torch(dist, dev)
        elif world > 1:
            raise SystemExit("--streamed runs with the in-launch peer exchange")
    else:
        if family in ("ordered_logistic", "categorical_logit"):
            # class-outcome models: standard-normal X, classes drawn uniformly
            from stan_b200.synth import shard_rows
            r0, r1 = shard_rows(N_total, rank, world)
            gen = torch.Generator(device=dev).manual_seed(20261017 + rank)
            X = torch.randn((K, r1 - r0), generator=gen, device=dev, dtype=torch.float64)
            y = torch.randint(1, args.classes + 1, (r1 - r0,), generator=gen, device=dev, dtype=torch.int32)
            grp = trials = None
            args.no_parity = True            # parity of these families: tests/test_class_models_gpu.py
        else:
            X, y, grp, trials, r0, r1 = make_shard_ex(torch, dev, family, N_total, K, G, rank, world, rows=rows)
        n_local = r1 - r0
        torch.cuda.synchronize()
        m = GLMModel(family, X.data_ptr(), y.data_ptr(), grp.data_ptr() if G else None, G, data_on_device=True,
                     N=n_local, K=K, ldx=n_local, device=local_rank, rank=rank, world=world, N_total=N_total,
                     trials=trials.data_ptr() if trials is not None else None, n_classes=args.classes
                     if family in ("ordered_logistic", "categorical_logit") else 0)
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
        # parity material (every rank): a row sample of this shard for the CPU checker, and -- memory permitting -- the
        # whole shard as a second, UNSHARDED handle for the additivity check
        par_rows = min(args.parity_rows, n_local) if not args.no_parity else 0
        par_sample = None
        if par_rows:
            par_sample = (X[:, :par_rows].cpu().numpy().T, y[:par_rows].cpu().numpy(),
                          grp[:par_rows].cpu().numpy() if G else None,
                          trials[:par_rows].cpu().numpy() if trials is not None else None)
        m_local = None
        free_b, _tot = torch.cuda.mem_get_info()
        if world > 1 and par_rows and free_b > 1.3 * 8 * n_local * (K + 3):
            m_local = GLMModel(family, X.data_ptr(), y.data_ptr(), grp.data_ptr() if G else None, G, data_on_device=True,
                               N=n_local, K=K, ldx=n_local, device=local_rank,
                               trials=trials.data_ptr() if trials is not None else None)
        del X, y, grp, trials
        torch.cuda.empty_cache()

    P = m.num_params_r()
    bytes_per_gradient = m.bytes_per_gradient()
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
    # the C entry point itself, called with preallocated host buffers (what a C++ caller does; the numpy / exception
    # plumbing of GLMModel.log_prob_grad costs a few microseconds per call and is not part of the boundary)
    import ctypes as C
    th = q0.copy()
    g_host, lp_host = np.empty(P), C.c_double()
    dp = C.POINTER(C.c_double)
    th_p, g_p, lp_p = th.ctypes.data_as(dp), g_host.ctypes.data_as(dp), C.byref(lp_host)
    call = m.L.b200glm_log_prob_grad
    for _ in range(max(3, args.warmup)):
        m.log_prob_grad(th)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        th[0] = q0[0] + 1e-6 * i
        rc = call(m.h, 0, th_p, 1, 1, lp_p, g_p)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if rc != 0 or not np.isfinite(lp_host.value):
        raise SystemExit(f"b200glm_log_prob_grad failed in the e2e loop (rc={rc})")
    e2e_ms = (t1 - t0) * 1000.0
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = args.steps / (e2e_ms / 1000.0)

    # ---- parity of what was just timed (every rank takes part; rank 0 reports) ----
    parity = None
    if par_sample is not None:
        parity = parity_record(torch, dist, dev, m, m_local, par_sample, family, G, K, rank, world, local_rank)
    if m_local is not None:
        m_local.close()
    # ---- four chains per pass (glm_multi_kernel): the same X, one batched leapfrog of 4 chain slots per
    #      step, device-resident; unsharded scalar-intercept handles with K <= 128 only ----
    four = None
    if world == 1 and G == 0 and K <= 128 and family in ("bernoulli_logit", "poisson_log", "normal_id") and not streamed:
        try:
            four = four_chains_record(args, torch, m, dev, bytes_per_gradient, q0)
        except Exception as e:                       # an extra: must not break the bench line
            four = {"error": str(e)[:200]}
    # ---- ESS/s inside NUTS at this configuration (every rank takes part; rank 0 reports) ----
    ess = None
    if args.ess_iters > 0 and args.config == 2:
        m.close()
        m = None
        torch.cuda.empty_cache()
        ess = ess_record(args, torch, dist, dev, rank, world, local_rank)
    read_gbs = None
    if rank == 0:
        from stan_b200 import _capi
        try:
            read_gbs, _ = _capi.measure_peaks(local_rank, read=True, dmma=False)
        except Exception:
            read_gbs = None

    if rank != 0:
        if m is not None:
            m.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, which = measured_peaks()
    bytes_per_launch = bytes_per_gradient               # this rank's shard: 8*N*K + 4*N
    achieved = bytes_per_launch / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": which,
                "kernel": (f"glm_class_kernel<{family}, {args.classes} classes>" if family in ("ordered_logistic", "categorical_logit")
                           else f"glm_wide_kernel<{family}>" if K > 256 else f"glm_fused_kernel<{family}>"), "algorithmic_bytes_per_launch": bytes_per_launch,
                "avg_launch_ms": ms_per_step}
    if read_gbs:
        # the kernel reads X once and writes nothing; `peak` above is the driver's read+write COPY figure, which a
        # read-only stream can exceed.  The same-process read-only figure is the tighter denominator.
        roofline["read_only_stream_gbs"] = read_gbs
        roofline["frac_of_read_only_stream"] = achieved / read_gbs
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            for ent in tj.get("entries", [tj]):
                if (ent.get("N") == n_local and ent.get("K") == K and ent.get("family", FAMILY) == family
                        and ent.get("G", 0) == G and ent.get("kernel", "").startswith(roofline["kernel"].split("<")[0])):
                    # dram__bytes_read.sum + dram__bytes_write.sum of THIS kernel at THIS shard size from a separate
                    # `ncu --set full` pass (profiles/README.md); null when no capture of this exact shape exists
                    roofline["traffic"] = ent.get("dram_bytes_per_launch")
                    roofline["traffic_source"] = ent.get("source")
        except Exception:
            pass

    cpu_baseline = None
    if sample is not None:
        cpu_baseline = cpu_baseline_leg(sample[0], sample[1], N_total, threads=1, evals=args.cpu_evals,
                                        family=family, group=sample[2], G=G, trials=sample[3],
                                        n_classes=args.classes if family in ("ordered_logistic", "categorical_logit") else 0)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N_total, K, family, G, args.config), "rows_total": N_total,
                   "rows_per_gpu": n_local, "cols": K, "groups": G,
                   **({"streamed_build_s": getattr(args, "build_seconds", None),
                       "hbm_gb_per_gpu": round(bytes_per_gradient / 1e9, 1)} if streamed else {}),
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
        "parity": parity,
        "four_chains": four,
        "ess": ess,
    }
    if ess is not None and world == 1 and not args.no_cpu_baseline:
        # BASELINE configs[0] in full, both arms measured in this run: the reference's CPU-runnable case (N=10k, K=20,
        # 4 chains x 1000+1000) through the same unmodified service on b200::glm_model and on the reference CPU model
        try:
            ess["config1_both_arms"] = ess_config1_record()
        except Exception as e:                                   # the bench line must not depend on this extra
            ess["config1_both_arms"] = {"error": str(e)[:200]}
    if ess and cpu_baseline and ess.get("b200"):
        # BASELINE.md section 5 step 3 at config 2: the CPU NUTS arm cannot run to an ESS inside a bench (1.8 s per
        # gradient per core); the same chains cost the same gradient evaluations on the CPU, so its ESS/s is the
        # chains' ESS per gradient evaluation times the CPU gradient rate measured above (projection, labelled)
        b = ess["b200"]
        ess["reference_cpu_projected"] = {
            "ess_min_per_s": b["ess_min"] / (b["grad_evals"] / cpu_baseline["value"]), "cores": cpu_baseline["cores"],
            "how": "ess_min of the chains above / (their gradient evaluations / cpu_baseline grad_evals/s); "
                   "one chain at a time on one core, as the b200 arm runs one chain at a time on the GPU(s)"}
    print(json.dumps(out))
    if m is not None:
        m.close()
    if world > 1:
        dist.destroy_process_group()


def build_streamed(torch, dist, dev, args, rank, world, local_rank):
    """BASELINE configs[4] at HBM scale: the design matrix is generated and handed to the backend chunk by chunk
    (B200GLM_FLAG_STREAMED / b200glm_append_rows), so it is resident ONCE -- as the panels the kernel streams.
    --rows 0: as many rows per GPU as fit (free memory minus head-room), the same on every rank.
    Synthetic data: one base chunk of standard normals per rank, chunk c = base * (1 + c / 1024) (cheap to produce,
    every row distinct); y ~ bernoulli_logit(0.3 + x . beta)."""
    from stan_b200 import GLMModel
    K, family = args.cols, args.family
    if family != "bernoulli_logit" or args.groups:
        raise SystemExit("--streamed is wired for the bernoulli_logit configuration without groups")
    chunk = 131_072
    if args.rows <= 0:
        free_b, _tot = torch.cuda.mem_get_info()
        t = torch.tensor([free_b], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        wr = 16 if K + 1 <= 512 else (8 if K + 1 <= 1024 else 4)          # wide_rows_for / wide_cps (glm_wide_kernel.cuh)
        cps = 32 // (wr // 4)
        cpad = (K + cps) // cps * cps if K > 256 else K + 1
        head = 6 * 8 * chunk * K + (6 << 30)                 # generator temporaries + workspaces + slack
        n_local = int((float(t.item()) - head) // (8 * cpad)) // chunk * chunk
    else:
        n_local = args.rows
    N_total = n_local * world
    m = GLMModel.streamed(family, n_local, K, device=local_rank, rank=rank, world=world, N_total=N_total)
    g = torch.Generator(device=dev).manual_seed(20261017 + rank)
    base = torch.randn((K, chunk), generator=g, device=dev, dtype=torch.float64)
    beta = torch.from_numpy(np.random.Generator(np.random.Philox(key=[20261017, 1])).standard_normal(K) / np.sqrt(K)).to(dev)
    sample = par_sample = None
    Xc = torch.empty_like(base)
    t0 = time.time()
    for c, r0 in enumerate(range(0, n_local, chunk)):
        n = min(chunk, n_local - r0)
        torch.mul(base, 1.0 + c / 1024.0, out=Xc)
        u = torch.rand(chunk, generator=g, device=dev, dtype=torch.float64)
        yc = (u < torch.sigmoid(0.3 + beta @ Xc)).to(torch.int32)
        if c == 0:
            pr = 0 if args.no_parity else min(args.parity_rows, 20_000, n)
            if pr:
                par_sample = (Xc[:, :pr].cpu().numpy().T, yc[:pr].cpu().numpy(), None, None)
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                ns = min(args.cpu_sample_rows, 100_000, n)
                sample = (Xc[:, :ns].cpu().numpy().T, yc[:ns].cpu().numpy(), None, None)
        torch.cuda.synchronize()
        m.append_rows(Xc.data_ptr(), yc.data_ptr(), None, n=n, ldx=chunk)
    m.finalize()
    args.build_seconds = time.time() - t0
    del base, Xc
    torch.cuda.empty_cache()
    return m, n_local, N_total, sample, par_sample


def parity_record(torch, dist, dev, m, m_local, sample, family, G, K, rank, world, local_rank):
    """Parity of the handles the bench just timed, checked in the same run:
    (1) kernel vs CPU checker (the compiled reference when present, else the C port) on a row sample of THIS
        rank's shard: the likelihood term and its partials through b200glm_glm_lpmf, relative error scaled as in
        tests/conftest.py; max over ranks;
    (2) world > 1: shard additivity -- the sharded handle's likelihood term (exchange inside the launch) against the
        sum over ranks of each shard evaluated by an unsharded handle on the same rows.
    The bar is north_star's 1e-10."""
    from stan_b200 import GLMModel
    from oracle.oracle import PortOracle, RefOracle
    Xs, ys, gs, ts = sample
    rng = np.random.default_rng(23)
    na = max(G, 1)
    alpha, beta = 0.1 * rng.standard_normal(na), 0.1 * rng.standard_normal(K)
    sigma = 1.3
    kw = {"trials": ts} if ts is not None else {}
    ms = GLMModel(family, np.asfortranarray(Xs), ys, gs, G, device=local_rank, **kw)
    lp, da, db, dsg = ms.glm_lpmf(alpha if G else alpha[0], beta, sigma)
    ms.close()
    if RefOracle.available():
        kind = "reference"
        lp_r, da_r, db_r, ds_r = RefOracle.glm_function(family, Xs, ys, alpha if G else alpha[0], beta, sigma,
                                                        group=gs, G=G, trials=ts)
        g, g_r = np.concatenate([da, db, [dsg]]), np.concatenate([da_r, db_r, [ds_r]])
    else:
        kind = "port"
        po = PortOracle(family, Xs, ys, gs, G, **kw)
        th = np.zeros(po.P)
        lp_r, g_r = po.log_prob_grad(th)
        mm = GLMModel(family, np.asfortranarray(Xs), ys, gs, G, device=local_rank, **kw)
        lp, g = mm.log_prob_grad(th)
        mm.close()
    sc = np.maximum(np.abs(g_r), np.max(np.abs(g_r)))
    sc = np.where(sc == 0, 1.0, sc)
    err_sample = max(abs(lp - lp_r) / max(abs(lp_r), 1e-300), float(np.max(np.abs(g - g_r) / sc)))
    err_add = None
    if world > 1:
        e = torch.tensor([err_sample], device=dev, dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        err_sample = float(e.item())
        if m_local is not None:
            lpl, dal, dbl, dsl = m_local.glm_lpmf(alpha if G else alpha[0], beta, sigma)
            part = torch.tensor(np.concatenate([[lpl], np.atleast_1d(dal), dbl]), device=dev, dtype=torch.float64)
            dist.all_reduce(part)
            lps, das, dbs, dss = m.glm_lpmf(alpha if G else alpha[0], beta, sigma)
            tot = np.concatenate([[lps], np.atleast_1d(das), dbs])
            ref = part.cpu().numpy()
            sc = np.maximum(np.abs(ref[1:]), np.max(np.abs(ref[1:])))
            err_add = max(abs(tot[0] - ref[0]) / abs(ref[0]), float(np.max(np.abs(tot[1:] - ref[1:]) / sc)))
            e = torch.tensor([err_add], device=dev, dtype=torch.float64)
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
            err_add = float(e.item())
    worst = max(err_sample, err_add or 0.0)
    return {"ok": bool(worst < 1e-10), "max_rel_err": worst, "tolerance": 1e-10,
            "sample_vs_cpu_checker": {"max_rel_err": err_sample, "rows_per_rank": int(Xs.shape[0]), "checker": kind,
                                      "what": "likelihood term + partials (b200glm_glm_lpmf) on a row sample of each rank's shard"},
            "shard_additivity": None if err_add is None else {
                "max_rel_err": err_add, "what": "sharded handle (in-launch exchange) vs sum over ranks of the "
                                                "same shards evaluated unsharded"}}


def four_chains_record(args, torch, m, dev, bytes_per_pass, q0):
    """Gradient evaluations/s with FOUR chains advanced per pass over X (b200glm_leapfrog_batched_async with 4 chain
    slots -> glm_multi_kernel + the batched prologue / reduce / epilogue launches), timed like `value`: CUDA events on
    the batch stream, args.warmup + args.steps steps, state resident on the device.  The roofline entry counts the
    ALGORITHMIC bytes of one pass (X once) -- four gradient evaluations ride on them."""
    C4 = 4
    m.batch_reserve(C4)
    rng = np.random.default_rng(12)
    q = q0[None, :] + 0.01 * rng.standard_normal((C4, q0.size))
    p = rng.standard_normal((C4, q0.size))
    lp, g, st = m.log_prob_grad_batched(q)
    if st.any():
        return {"error": "domain error at the starting points"}
    m.set_state_batched(q, p, -g, -lp)
    stream = torch.cuda.ExternalStream(m.batch_stream_ptr(), device=dev)
    for _ in range(args.warmup):
        m.leapfrog_batched_async(C4, 1e-4)
    torch.cuda.synchronize()
    l0 = m.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        m.leapfrog_batched_async(C4, 1e-4)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    peak, which = measured_peaks()
    achieved = bytes_per_pass / (ms * 1e-3) / 1e9
    return {"chains": C4, "value": C4 * 1000.0 / ms, "unit": UNIT, "ms_per_step": ms,
            "gpu_launches": int(m.launch_count() - l0),
            "step": "one batched leapfrog of 4 chain slots: batched_begin + glm_multi_kernel (one pass over X for the four "
                    "chains) + batched_reduce + batched_finish, device-resident state",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peak["hbm_gbs"], "peak_source": which,
                         "kernel": "glm_multi_kernel (4 chains per pass)",
                         "algorithmic_bytes_per_launch": int(bytes_per_pass), "avg_launch_ms": ms,
                         "note": "time of the whole 4-launch step over the bytes of ONE pass over X"}}


def ess_record(args, torch, dist, dev, rank, world, local_rank):
    """min ESS/s inside NUTS at the bench configuration: the reference's unmodified hmc_nuts_diag_e_adapt
    (libb200stan.so) on b200::glm_model, `--ess-chains` chains run one after the other (every rank runs the same
    deterministic host code on the same seeds; a chain's leapfrog is one launch per rank), `--ess-iters` warm-up +
    `--ess-iters` sampling iterations each.  ESS by stan::analyze::ess (compiled into the shim)."""
    from stan_b200 import stan_service
    from stan_b200.synth import make_shard_ex
    if not stan_service.available():
        return {"unavailable": "stan_b200/lib/libb200stan.so not built (needs the reference headers at build time)"}
    N_total, K = args.rows, args.cols
    X, y, _, _, r0, r1 = make_shard_ex(torch, dev, args.family, N_total, K, 0, rank, world)
    torch.cuda.synchronize()
    sm = stan_service.StanGLM(args.family, X.data_ptr(), y.data_ptr(), device=local_rank, n_slots=1, rank=rank,
                              world=world, N_total=N_total, data_on_device=True, N=r1 - r0, K=K, ldx=r1 - r0)
    del X, y
    torch.cuda.empty_cache()
    if world > 1:
        sm.connect_peers_torch(dist, dev)
        dist.barrier()
    it = args.ess_iters
    res = sm.nuts(num_chains=args.ess_chains, seed=4711, num_warmup=it, num_samples=it, delta=0.8, num_threads=1)
    wall = res["wall"]
    if world > 1:
        t = torch.tensor([wall], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    # second arm: the same chains (same seeds) through b200::hmc_nuts_diag_e_adapt_device -- the chains advance
    # TOGETHER, one pass over X per leapfrog of all of them (few-chain FMA kernel for <= 8 lanes, fp64 DMMA kernels
    # above), tree building and adaptation on the device.  On row shards every rank passes over ITS rows
    # for all chains and one NCCL all-reduce per round combines the (K + 2) x chains partial sums.
    dres, derr = None, None
    if world == 1 or not args.no_ess_device_sharded:
        try:
            if world > 1:
                sm.comm_init_torch(dist, dev)
            dres = sm.nuts_device(num_chains=args.ess_chains, seed=4711, num_warmup=it, num_samples=it, delta=0.8)
        except Exception as e:   # a shape outside the batched kernel (K > 208): reported, not fatal (deterministic: every
            derr = str(e)[:200]  # rank takes the same branch)
        if dres is not None and world > 1:
            t = torch.tensor([dres["wall"]], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dres["wall"] = float(t.item())
    sm.close()
    if rank != 0:
        return None

    def arm(r, wall_s, n_grad):
        d = r["draws"]
        P = d.shape[2] - 7
        ess = [stan_service.diagnostic("ess", d[:, :, 7 + k].T) for k in range(P)]
        rhat = [stan_service.diagnostic("rhat", d[:, :, 7 + k].T) for k in range(P)]
        return {"chains": args.ess_chains, "iters": f"{it}+{it}", "wall_s": wall_s, "grad_evals": n_grad,
                "grad_evals_per_s": n_grad / wall_s, "ess_min": float(np.min(ess)), "ess_median": float(np.median(ess)),
                "ess_min_per_s": float(np.min(ess)) / wall_s, "rhat_max": float(np.max(rhat)),
                "mean_treedepth": float(d[:, :, 3].mean()), "divergent": int(d[:, :, 5].sum()),
                "stepsize": [float(v) for v in r["stepsize"]]}
    d = res["draws"]
    out = {"b200": arm(res, wall, float(d[:, :, 4].sum() + res["warm_leapfrogs"].sum() + 2 * it * args.ess_chains)),
           "service": "stan::services::sample::hmc_nuts_diag_e_adapt (unmodified) on b200::glm_model, chains sequential"}
    if dres is not None:
        out["b200_device_driver"] = dict(
            arm(dres, dres["wall"], float(dres["lanes"])), rounds=dres["rounds"],
            service="b200::hmc_nuts_diag_e_adapt_device: same seeds, the chains advance together (one pass over X per "
                    "leapfrog of all chains), NUTS transition and adaptation on the device")
        out["b200_device_driver"]["first_draws_max_abs_diff_vs_service"] = float(
            np.max(np.abs(dres["warmup_draws"][:, :3, 7:] - res["warmup_draws"][:, :3, 7:])))
    elif derr:
        out["b200_device_driver"] = {"unavailable": derr}
    return out


def ess_config1_record():
    """bernoulli_logit N=10k K=20, NUTS diag_e 4 chains 1000+1000, same seeds on both arms: wall time, gradient
    evaluations/s and min ESS/s of the GPU arm (libb200stan.so) and of the reference CPU arm (oracle/_ref, 4 threads),
    and the largest posterior-mean z-score between them."""
    from stan_b200 import make_glm_data, stan_service
    from oracle.oracle import RefOracle
    if not (stan_service.available() and RefOracle.available()):
        return {"unavailable": "needs libb200stan.so and oracle/_ref"}
    d = make_glm_data("bernoulli_logit", 10_000, 20)
    kw = dict(num_chains=4, seed=4711, num_warmup=1000, num_samples=1000, delta=0.8, num_threads=4)
    m = stan_service.StanGLM("bernoulli_logit", d["X"], d["y"], n_slots=8)
    dev = m.nuts(**kw)
    m.close()
    ref = RefOracle("bernoulli_logit", d["X"], d["y"]).nuts(**kw)

    def arm(r):
        dr = r["draws"]
        ess = [stan_service.diagnostic("ess", dr[:, :, 7 + k].T) for k in range(dr.shape[2] - 7)]
        n_grad = float(dr[:, :, 4].sum() + r["warm_leapfrogs"].sum() + dr.shape[0] * 2000)
        return {"wall_s": r["wall"], "grad_evals_per_s": n_grad / r["wall"], "ess_min": float(np.min(ess)),
                "ess_min_per_s": float(np.min(ess)) / r["wall"], "divergent": int(dr[:, :, 5].sum())}
    zs = []
    for k in range(dev["draws"].shape[2] - 7):
        a, b = dev["draws"][:, :, 7 + k].T, ref["draws"][:, :, 7 + k].T
        zs.append(abs(a.mean() - b.mean()) / np.hypot(stan_service.diagnostic("mcse_mean", a),
                                                       stan_service.diagnostic("mcse_mean", b)))
    out = {"workload": "bernoulli_logit_glm N=10000 K=20, NUTS diag_e 4 chains 1000+1000 (BASELINE configs[0])",
           "b200": arm(dev), "reference_cpu": dict(arm(ref), threads=4), "posterior_mean_max_z": float(max(zs))}
    out["ess_per_s_ratio"] = out["b200"]["ess_min_per_s"] / out["reference_cpu"]["ess_min_per_s"]
    return out


def run_b200_batched(args):
    """BASELINE configs[2]: normal_id_glm N=1M K=200, 1024 batched chains on one B200.  A step = one
    batched leapfrog (every chain advances once = `chains` gradient evaluations): fp64 DMMA GEMM pair."""
    import torch
    import torch.distributed as dist
    from stan_b200 import GLMModel
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, K, C_total = args.rows, args.cols, args.chains
    # N GPUs: the CHAINS are what shards (the way the reference's multi-chain service parallelises,
    # hmc_nuts_diag_e_adapt.hpp:387-401): every rank holds all of X (1.6 GB) and --chains / N of the chains; no
    # data-path collective.  Strong scaling: the job (N x K x chains) is fixed.
    C = C_total // world + (1 if rank < C_total % world else 0)
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
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        m.leapfrog_batched_async(C, eps)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:   # max over ranks, measured on the devices
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    launches = m.launch_count() - launches0
    clk = clocks.stop(t_wall0, t_wall1)
    value = C_total * 1000.0 / ms_per_step
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
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = C_total * ne / e2e_s
    flops = 4.0 * N * K * C                     # this rank's launch (the roofline is per kernel launch)
    # fp64 tensor-pipe peak measured NOW, in this process and under the clocks of this run (MEASURED_PEAKS.json has
    # no fp64 entry): register-resident mma.sync.m8n8k4.f64 loop, b200glm_measure_peaks (stan_b200/csrc/measure.cuh)
    from stan_b200 import _capi
    clk2 = ClockSampler(local_rank)
    clk2.start()
    tp0 = time.time()
    _, peak = _capi.measure_peaks(local_rank, read=False, dmma=True)
    peak_clocks = clk2.stop(tp0, time.time())
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    cpu_baseline = None
    if rank != 0:
        m.close()
        dist.destroy_process_group()
        return
    if sample is not None:
        cpu_baseline = cpu_baseline_leg(sample[0], sample[1], N, threads=1, evals=args.cpu_evals, family="normal_id")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"normal_id_glm N={N} K={K} fp64, {C_total} batched chains (BASELINE configs[2])",
                   "rows_total": N, "cols": K, "chains": C_total, "chains_per_gpu": C,
                   "sharding": "single GPU" if world == 1 else f"chains x{world} (X replicated on every GPU, no collective)",
                   "l2": f"X {8e-9 * N * K:.2f} GB >> 126 MB L2, no flush needed",
                   "step": "one batched leapfrog: begin + fused DMMA GEMM pair + slice reduce + finish (device-resident state)"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * P * C_total, "d2h_bytes_per_step": 8 * (P + 2) * C_total,
                "call": "b200glm_log_prob_grad_batched(host thetas) -> host lp, grad"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": "fp64 DMMA register loop measured in this run (b200glm_measure_peaks)",
                     "peak_clocks": peak_clocks,
                     "kernel": "glm_batched_kernel<normal_id,13>", "algorithmic_flops_per_launch": flops,
                     "avg_launch_ms": ms_per_step},
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(out))
    m.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------
def cpu_baseline_leg(Xs, ys, N_total, threads=1, evals=5, family=FAMILY, group=None, G=0, trials=None, n_classes=0):
    from oracle.oracle import PortOracle, RefOracle
    cls = RefOracle if RefOracle.available() else PortOracle
    kw = {"trials": trials} if trials is not None else {}
    if n_classes:
        kw["n_classes"] = n_classes
    orc = cls(family, Xs, ys, group, G, **kw)
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
    budget_s = 240.0
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
          "sample": (f"all {args.rows} rows (the full configuration)" if ns == args.rows else
                     f"{ns} of {args.rows} rows, rate scaled by rows ({ns}/{args.rows})")
                    + f"; each step = {threads} concurrent chains (threads) each doing one stan::model::log_prob_grad",
          "same_config": bool(ns == args.rows)}
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
    ap.add_argument("--streamed", action="store_true",
                    help="config 5: build the handle chunk by chunk (X resident once); --rows 0 = as many rows per GPU as fit")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: split the rows in proportion to each GPU's measured copy bandwidth (default: equal)")
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--classes", type=int, default=4, help="--family ordered_logistic | categorical_logit: outcome classes")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the P+2 likelihood partials are summed over ranks")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--cols", type=int, default=None)
    ap.add_argument("--cpu-sample-rows", type=int, default=None,
                    help="rows of the CPU legs' sample (default: 1M for cpu_baseline; ALL rows for --impl reference)")
    ap.add_argument("--parity-rows", type=int, default=200_000, help="rows per rank checked against the CPU checker")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--ess-iters", type=int, default=100,
                    help="config 2: warm-up = sampling iterations of the in-bench NUTS run (0 skips it)")
    ap.add_argument("--ess-chains", type=int, default=4)
    ap.add_argument("--no-ess-device-sharded", action="store_true",
                    help="N > 1: skip the device-side NUTS arm on row shards (chains advance together, one NCCL "
                         "all-reduce per round); the unmodified service arm always runs")
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
        args.cpu_sample_rows = args.cpu_sample_rows if args.cpu_sample_rows is not None else 1_000_000
        args.steps = args.steps if args.steps is not None else 20
        args.warmup = max(3, args.warmup if args.warmup is not None else 3)
        return run_b200_batched(args)
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        args.cpu_sample_rows = args.cpu_sample_rows if args.cpu_sample_rows is not None else args.rows
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 200
        args.warmup = max(3, args.warmup if args.warmup is not None else 10)
        args.cpu_sample_rows = args.cpu_sample_rows if args.cpu_sample_rows is not None else 1_000_000
        run_b200(args)


if __name__ == "__main__":
    main()
