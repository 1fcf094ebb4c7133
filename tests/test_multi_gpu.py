"""Device-side multi-GPU parity: runs tests/multi_gpu_check.py under torchrun when the box has >= 2 GPUs (the halo
exchange over NCCL/NVLink must make box boundaries invisible: bit-identical to the single-box run in the exact
build).  With one GPU the test is skipped; the host logic is covered on CPU by tests/test_level_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["ss", "fe"])
def test_two_gpu_level_matches_single_box(model):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--size", "40", "--model", model]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bit-identical = True" in r.stdout
