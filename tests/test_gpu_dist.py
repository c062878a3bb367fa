"""Multi-GPU test of the path's one exchange step: batch-summed gradients all-reduced with NCCL across the GPUs of
one box, checked on the GPUs against the per-problem gradients of every rank.  Needs >= 2 devices (skipped otherwise; run
with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_allreduce_of_batch_summed_gradients():
    import idoc_b200 as ib
    ndev = ib._lib.device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ndev < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29653", os.path.join(ROOT, "tests", "_dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "allreduce verified" in r.stdout
