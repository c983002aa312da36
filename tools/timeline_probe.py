"""Where a gradient launch spends its time: per-phase %globaltimer stamps of the narrow kernel and its tail
(b200glm_timeline_*).  One GPU:  python tools/timeline_probe.py --rows 1250000
N GPUs (row shards, in-kernel peer exchange):  torchrun --nproc-per-node N tools/timeline_probe.py --rows 10000000
Prints, per rank, medians over `--reps` launches (ns relative to the first CTA's entry) and the step time by CUDA
events with the stamps switched off, so the table can be checked against the un-instrumented launch."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stan_b200 import GLMModel  # noqa: E402
from stan_b200.synth import make_shard_ex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_250_000, help="rows over all ranks")
    ap.add_argument("--cols", type=int, default=100)
    ap.add_argument("--family", default="bernoulli_logit")
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    X, y, grp, trials, r0, r1 = make_shard_ex(torch, dev, a.family, a.rows, a.cols, a.groups, rank, world)
    m = GLMModel(a.family, X.data_ptr(), y.data_ptr(), grp.data_ptr() if a.groups else None, a.groups,
                 data_on_device=True, N=r1 - r0, K=a.cols, ldx=r1 - r0, device=local, rank=rank, world=world,
                 N_total=a.rows, trials=trials.data_ptr() if trials is not None else None)
    if world > 1:
        m.connect_peers_torch(dist, dev)
    del X, y
    P = m.num_params_r()
    rng = np.random.default_rng(11)
    q0, p0 = 0.05 * rng.standard_normal(P), rng.standard_normal(P)
    lp0, g0 = m.log_prob_grad(q0)
    m.set_state(q0, p0, -g0, -lp0)
    stream = torch.cuda.ExternalStream(m.stream_ptr(0), device=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record(stream)
        for _ in range(n):
            m.leapfrog_async(1e-4)
        e1.record(stream)
        sync()
        return e0.elapsed_time(e1) / n * 1e3   # us

    for _ in range(20):
        m.leapfrog_async(1e-4)
    us_plain = timed(a.reps)
    m.timeline_enable(True)
    us_stamped = timed(a.reps)
    # back-to-back launches overwrite the buffer: sample single launches, each after a run-up of 3 so that the
    # programmatic-dependent-launch overlap with the previous launch is what a steady-state step sees
    rows = []
    for _ in range(a.reps):
        sync()
        for _ in range(4):
            m.leapfrog_async(1e-4)
        m.sync()
        rows.append(m.timeline_read())
    T = np.stack(rows).astype(np.int64)          # reps x (grid + 2) x 16
    grid = T.shape[1] - 2
    cta, tail, fine = T[:, :grid, :], T[:, grid, :], T[:, grid + 1, :]
    t0 = cta[:, :, 0].min(axis=1)                # first CTA entry of the launch
    med = lambda x: float(np.median(x))          # noqa: E731
    rel = lambda x: x - t0[:, None] if x.ndim == 2 else x - t0   # noqa: E731
    out = {
        "rank": rank, "world": world, "rows_local": r1 - r0, "cols": a.cols, "grid": grid,
        "us_per_step_events_plain": us_plain, "us_per_step_events_stamped": us_stamped,
        "globaltimer_resolution_ns": int(np.min(np.diff(np.unique(T[0, :, :6].ravel()))[np.diff(np.unique(T[0, :, :6].ravel())) > 0])),
        "ns_since_first_cta_entry": {
            "cta_entry_last": med(rel(cta[:, :, 0]).max(axis=1)),
            "pdl_wait_over_median_cta": med(np.median(rel(cta[:, :, 1]), axis=1)),
            "theta_staged_median_cta": med(np.median(rel(cta[:, :, 2]), axis=1)),
            "first_panel_landed_median_cta": med(np.median(rel(cta[:, :, 3]), axis=1)),
            "last_panel_consumed_first_cta": med(rel(cta[:, :, 4]).min(axis=1)),
            "last_panel_consumed_median_cta": med(np.median(rel(cta[:, :, 4]), axis=1)),
            "last_panel_consumed_last_cta": med(rel(cta[:, :, 4]).max(axis=1)),
            "partial_written_last_cta": med(rel(cta[:, :, 5]).max(axis=1)),
            "tail_ticket_won": med(rel(tail[:, 0])),
            "tail_grid_sum_done": med(rel(tail[:, 1])),
            "tail_peer_exchange_done": med(rel(tail[:, 2])),
            "tail_finish_done": med(rel(tail[:, 3])),
            "fine_fence_after_ticket": med(rel(fine[:, 0])),
            "fine_first_rows_loaded": med(rel(fine[:, 1])),
            "fine_sums_written": med(rel(fine[:, 2])),
            "fine_finish_value_done": med(rel(fine[:, 3])),
            "fine_finish_gradient_written": med(rel(fine[:, 4])),
            "fine_cold_pass_done": med(rel(fine[:, 5])) if os.environ.get("B200GLM_TL_REPEAT") == "1" else None,
        },
    }
    d = out["ns_since_first_cta_entry"]
    out["phases_us"] = {
        "fill: entry -> first panel landed": (d["first_panel_landed_median_cta"]) / 1e3,
        "stream: first panel -> median CTA done": (d["last_panel_consumed_median_cta"] - d["first_panel_landed_median_cta"]) / 1e3,
        "skew: median CTA done -> last CTA done": (d["last_panel_consumed_last_cta"] - d["last_panel_consumed_median_cta"]) / 1e3,
        "cta reduce + partial write + ticket": (d["tail_ticket_won"] - d["last_panel_consumed_last_cta"]) / 1e3,
        "grid sum (last CTA)": (d["tail_grid_sum_done"] - d["tail_ticket_won"]) / 1e3,
        "peer exchange": (d["tail_peer_exchange_done"] - d["tail_grid_sum_done"]) / 1e3,
        "finish (epilogue + leapfrog tail)": (d["tail_finish_done"] - d["tail_peer_exchange_done"]) / 1e3,
        "launch total (first entry -> finish)": d["tail_finish_done"] / 1e3,
    }
    out["tail_fine_us"] = {
        "ticket won -> acquire fence": (d["fine_fence_after_ticket"] - d["tail_ticket_won"]) / 1e3,
        "-> first batch of partial rows loaded": (d["fine_first_rows_loaded"] - d["fine_fence_after_ticket"]) / 1e3,
        "-> sums written": (d["fine_sums_written"] - d["fine_first_rows_loaded"]) / 1e3,
        "-> fence + barrier (grid sum done)": (d["tail_grid_sum_done"] - d["fine_sums_written"]) / 1e3,
        "epilogue: -> block sums + value": (d["fine_finish_value_done"] - d["tail_peer_exchange_done"]) / 1e3,
        "epilogue: -> gradient + leapfrog tail written": (d["fine_finish_gradient_written"] - d["fine_finish_value_done"]) / 1e3,
        "epilogue: -> published": (d["tail_finish_done"] - d["fine_finish_gradient_written"]) / 1e3,
    }
    hbm_us = m.bytes_per_gradient() / 7.2e12 * 1e6
    out["shard_hbm_time_us_at_7.2TBps"] = hbm_us
    line = json.dumps(out)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, line)
        lines = gathered
    else:
        lines = [line]
    if rank == 0:
        for ln in lines:
            print(ln)
        if a.out:
            with open(a.out, "w") as f:
                f.write("\n".join(lines) + "\n")
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
