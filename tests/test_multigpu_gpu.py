"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): tools/mg_check.py under torchrun — the partitioned
2-rank run (Morton-range partition, NVLink peer pulls inside K1/K2/K3 and the interface pre-pass) must be bit-identical
to the single-GPU run on the synthetic box and on the full-featured two-level case."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_run_is_bit_identical_to_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tools", "mg_check.py")], capture_output=True, text=True, timeout=600)
    assert "MG_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
