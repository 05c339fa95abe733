"""N-GPU vs 1-GPU equivalence of the data-parallel path (needs >= 2 GPUs: skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_gpu_syncbn_equals_single_gpu(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dp_equiv_worker.py"), precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dp-equivalence ok rank 0" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
