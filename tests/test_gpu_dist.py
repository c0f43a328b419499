"""GPU parity of the multi-GPU path: partitioned wave / shallow-water steppers (one rank per GPU, halo exchange and dot
products through peer memory) against the single-GPU run.  Needs at least 2 GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n,p", [(6, 1), (5, 2), (8, 0)])
def test_partitioned_steppers_match_single_gpu(n, p):
    world = min(_ngpus(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "dist_worker.py"), str(n), str(p)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
