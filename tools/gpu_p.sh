#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_abi_and_oracle_units.py tests/test_k1_single_level_gpu.py tests/test_large_sizes_gpu.py tests/test_virtual_ranks_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 900 python tools/ab_box.py --nb 64 --steps 40 --repeat 2 "base|strict|" "base|fast|" > gpurun_out/p_ab.log 2>&1; echo "ab exit $?" >> gpurun_out/p_ab.log
grep -E "^AB|exit|Error" gpurun_out/p_ab.log | cut -c1-360
