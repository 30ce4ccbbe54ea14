#!/bin/bash
mkdir -p gpurun_out
# is compute-sanitizer usable on this pool now?  (round 1: refused)  two-level feature case, 2 virtual ranks + single context, both FP modes
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_virtual_ranks_gpu.py -m gpu -q -x -p no:cacheprovider -k "two_level and 2-None-False" > gpurun_out/n_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/n_memcheck.log
tail -12 gpurun_out/n_memcheck.log | cut -c1-300
