#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_full_size_cases_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest.log
grep -E "passed|failed|Error|assert|exit" gpurun_out/m_pytest.log | cut -c1-600 | head -20
timeout 900 python tools/n2_full_size.py bunny_fine > gpurun_out/m_n2_full_size.log 2>&1; tail -5 gpurun_out/m_n2_full_size.log
