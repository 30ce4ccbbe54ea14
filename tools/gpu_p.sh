#!/bin/bash
timeout 300 python -m pytest tests/test_large_sizes_gpu.py tests/test_zz_output_gather.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
