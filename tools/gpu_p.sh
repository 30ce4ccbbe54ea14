#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_virtual_ranks_gpu.py tests/test_checkpoint_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4
