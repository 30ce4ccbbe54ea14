#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_k1_features_gpu.py tests/test_bunny_small_gpu.py tests/test_full_size_cases_gpu.py tests/test_cases_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4
timeout 1200 python tools/run_case_mg.py bunny 8 --fp-mode strict --uniform-start --variant "verbose=0" --variant "verbose=0" > gpurun_out/p_bunny.log 2>&1
grep -E "RESULT|exit|Error" gpurun_out/p_bunny.log | cut -c1-200
timeout 1200 python tools/run_case_mg.py wing5 8 --fp-mode strict --uniform-start --variant "verbose=0" --variant "verbose=0" > gpurun_out/p_wing.log 2>&1
grep -E "RESULT|exit|Error" gpurun_out/p_wing.log | cut -c1-200
