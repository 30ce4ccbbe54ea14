#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_block_order_gpu.py tests/test_multigpu_gpu.py tests/test_virtual_ranks_gpu.py tests/test_checkpoint_gpu.py tests/test_zz_output_gather.py tests/test_full_size_cases_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/s_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/s_pytest2.log
tail -4 gpurun_out/s_pytest2.log | cut -c1-300
