"""GPU (>= 2 devices): the slab-partitioned solver against the single-GPU one,
bit for bit, by running tests/dist_check.py under torchrun.  Skipped on a
one-GPU box (NCCL refuses two ranks on one device)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu(mgb):
    import ctypes as C
    n = C.c_int()
    mgb.load_library().mgb_device_count(n)
    return n.value


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_matches_single_gpu(mgb, world):
    if _ngpu(mgb) < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "dist_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("DIST_CHECK_OK") == world
