"""Multi-GPU form of one controller step: needs >= 2 GPUs on the box, skipped otherwise.  The same
split on a single GPU (no NCCL) is covered by test_gpu_parity.py::test_split_step_equals_single_handle and the
host-side exchange logic by the world-size-2 gloo test in test_host.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_split_step_over_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run with `gpurun --gpus 2`); the 1-GPU form of the split is tested in test_gpu_parity.py")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29741", os.path.join(ROOT, "tests", "multi_gpu_split_step.py")]  # fmt: skip
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "OK" in r.stdout
    print(r.stdout)
