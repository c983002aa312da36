"""GPU suite, needs >= 2 GPUs (run with `gpurun --gpus 2`): row-sharded model == the unsharded oracle and
replicated theta is bitwise identical on every rank, for both transports of the likelihood partials:
in-kernel peer mailboxes over NVLink (one launch per gradient) and one NCCL all-reduce per gradient."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 8])
def test_row_sharded_matches_oracle(world, transport):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29733" if transport == "peer" else "29734",
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, MGPU_TRANSPORT=transport))
    assert "MGPU-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
