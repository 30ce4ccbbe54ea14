#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_bunny_small_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider -k deeper > gpurun_out/f2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f2_pytest.log
grep -E "bunny res|passed|failed|Error|assert" gpurun_out/f2_pytest.log | cut -c1-2500 | head -40
