"""x-y decomposition over several GPUs with the NCCL halo exchange must reproduce the single-GPU run
(needs >= 2 GPUs on the box; skipped otherwise -- the strip logic itself is covered on CPU by test_cpu_halo_gloo)."""
import os
import subprocess
import sys

import pytest

from cgfd3d_b200 import solver

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("split,medium", [("x", "iso"), ("y", "iso"), ("x", "visco"), ("y", "vti")])
def test_two_ranks_match_one(split, medium):
    n = solver.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs, found %d" % n)
    port = 29631 + ["x", "y"].index(split) + 2 * ["iso", "vti", "visco"].index(medium)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "multi_gpu_check.py"), split, medium]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTI_GPU_CHECK" in p.stdout and '"ok": true' in p.stdout, p.stdout[-2000:]
