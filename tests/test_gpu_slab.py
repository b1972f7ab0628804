"""Multi-GPU parity of the y-slab decomposition: spawns tools/slab_check.py under torchrun on 2 GPUs
(skipped on a single-GPU box; tools/slab_check.py is also run directly with `gpurun --gpus N`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_slab_parity_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "slab_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "SLAB CHECK PASSED" in r.stdout, r.stderr[-3000:]
