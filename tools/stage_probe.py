"""Debug probe: does a (rows, cols, groups, ring depth) combination run?  python tools/stage_probe.py ROWS COLS G"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import GLMModel
from stan_b200.synth import make_shard_ex
N, K, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fam = sys.argv[4] if len(sys.argv) > 4 else "poisson_log"
dev = torch.device("cuda", 0)
X, y, grp, tr, r0, r1 = make_shard_ex(torch, dev, fam, N, K, G, 0, 1)
for S in [int(v) for v in os.environ.get("PROBE_STAGES", "8").split(",")]:
    os.environ["B200GLM_STAGES"] = str(S)
    try:
        m = GLMModel(fam, X.data_ptr(), y.data_ptr(), grp.data_ptr() if G else None, G, data_on_device=True, N=N, K=K, ldx=N)
        th = 0.05 * np.random.default_rng(1).standard_normal(m.P)
        lp, g = m.log_prob_grad(th)
        m.set_state(th, th, -g, -lp)
        t0 = time.time()
        for _ in range(30):
            m.leapfrog_async(1e-4)
        m.sync()
        dt = (time.time() - t0) / 30
        lp2, g2 = m.log_prob_grad(th)
        print(f"N={N} K={K} G={G} S={S}: ok  {dt*1e3:.3f} ms/step  {m.bytes_per_gradient()/dt/1e12:.2f} TB/s  same={lp==lp2 and np.array_equal(g,g2)}", flush=True)
        m.close()
    except Exception as e:
        print(f"N={N} K={K} G={G} S={S}: FAIL {e}", flush=True)
        break
