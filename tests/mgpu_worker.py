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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = []
    for fam, N, K, G in [("bernoulli_logit", 200_003, 100, 0), ("poisson_log", 150_000, 50, 1000),
                         ("normal_id", 90_001, 33, 0)]:
        d = make_glm_data(fam, N, K, G)                     # every rank builds the same global data
        r0, r1 = shard_rows(N, rank, world)
        grp = None if d["group"] is None else d["group"][r0:r1]
        m = GLMModel(fam, d["X"][r0:r1], d["y"][r0:r1], grp, G, device=local, rank=rank, world=world, N_total=N)
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
            po = PortOracle(fam, d["X"], d["y"], d["group"], G)
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
    dist.barrier()
    if rank == 0:
        print("MGPU-OK" if not failures else "MGPU-FAIL " + "; ".join(failures), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
