"""Multi-GPU parity, one process per rank: tools/mg_check.py under torchrun (2 GPUs when the box has them, otherwise both ranks on cuda:0) — the partitioned
2-rank run (Morton-range partition, NVLink peer pulls inside K1/K2/K3 and the interface pre-pass) must be bit-identical
to the single-GPU run on the synthetic box and on the full-featured two-level case."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_run_is_bit_identical_to_one_gpu():
    # with fewer than 2 GPUs the two processes share cuda:0 (gloo plumbing, CUDA-IPC peer mappings on one device, host barrier)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tools", "mg_check.py")], capture_output=True, text=True, timeout=600)
    assert "MG_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
