#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_launch_variants_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -8
