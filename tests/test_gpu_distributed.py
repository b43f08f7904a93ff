"""Multi-GPU parity under pytest -m gpu: every distributed result (sort, sort_by_key, scans, reductions) must equal the
single-GPU result bit for bit (floats: within tolerance).  One process per GPU under torchrun on 127.0.0.1; skipped when
fewer than two GPUs are visible.  The CUDA exchange kernel really crosses devices here (the gloo tests in
test_distributed_cpu.py only cover the host-side protocol)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("plan", ["peer-scatter", "nccl-all-to-all"])
def test_distributed_results_equal_single_gpu(plan):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 visible GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    env = dict(os.environ, BCB_DIST_PEER="1" if plan == "peer-scatter" else "0")
    port = 29600 + (os.getpid() % 200) + (0 if plan == "peer-scatter" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, tail
    assert "MISMATCH" not in r.stdout, tail
