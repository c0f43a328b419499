"""GPU parity of the multi-GPU path: partitioned wave / shallow-water steppers (one rank per GPU, halo exchange and dot
products through peer memory) against the single-GPU run.  On a box with one GPU the ranks share it (time-sliced processes,
gloo for the set-up exchange); nothing is skipped."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _run(world, n, p, share):
    env = dict(os.environ)
    if share:
        # one GPU on the box: the ranks are separate processes time-slicing GPU 0 (CUDA IPC maps a window of another
        # process on the same device just as well; gloo replaces NCCL, which refuses two ranks on one device).  The
        # partition, ghost layers, halo exchange and peer-memory all-reduce are the real ones; only the wire is not NVLink.
        env.update(GG_DIST_BACKEND="gloo", GG_DIST_SHARE_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "dist_worker.py"), str(n), str(p)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("n,p", [(6, 1), (5, 2), (8, 0)])
def test_partitioned_steppers_match_single_gpu(n, p):
    ng = _ngpus()
    if ng >= 2:
        _run(min(ng, 4), n, p, share=False)
    else:
        _run(2, n, p, share=True)


def test_partitioned_code_path_on_one_rank():
    """GG_FORCE_MULTI=1: the partitioned solver path (pushes, flag rounds, unpacks) with a world of one rank."""
    env = dict(os.environ, GG_FORCE_MULTI="1", GG_DIST_BACKEND="gloo", GG_DIST_SHARE_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1",
           "--master-port", "29618", os.path.join(ROOT, "tests", "dist_worker.py"), "4", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
