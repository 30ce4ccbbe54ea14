#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flood_fill_gpu.py tests/test_domain_gpu.py -m gpu -q -x -s --tb=short -p no:cacheprovider > gpurun_out/u_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/u_pytest.log
grep -E "passed|failed|Error|assert|exit|build" gpurun_out/u_pytest.log | cut -c1-300 | head -30
timeout 900 python tools/n2_full_size.py bunny_fine > gpurun_out/u_n2_full_size.log 2>&1; tail -5 gpurun_out/u_n2_full_size.log
