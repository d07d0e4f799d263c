"""The NCCL combine of per-rank partial framebuffers (lb_filter_reduce, lb_filter_reduce_scatter + lb_imager_resolve_gather)
against a single-GPU run over all samples.  Needs >= 2 visible GPUs: the test launches tests/multi_gpu_check.py under
torchrun with two ranks and is skipped on a one-GPU box (the driver's 1-GPU test tier)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_two_rank_reduce_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540, cwd=ROOT)
    assert r.returncode == 0 and "multi_gpu_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
