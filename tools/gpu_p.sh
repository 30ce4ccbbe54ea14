#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_options.py "face_persist=0" "face_persist=4" > gpurun_out/p_check.log 2>&1; echo "check exit $?" >> gpurun_out/p_check.log
grep -E "CHECK|exit|Error" gpurun_out/p_check.log | cut -c1-300
timeout 300 python -m pytest tests/test_large_sizes_gpu.py tests/test_graph_replay_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 900 python tools/ab_box.py --nb 64 --steps 40 --repeat 3 "fp0|strict|face_persist=0" "fp1|strict|face_persist=1" "fp2|strict|face_persist=2" "fp4|strict|face_persist=4" "fp0|fast|face_persist=0" "fp2|fast|face_persist=2" "fp4|fast|face_persist=4" > gpurun_out/p_ab.log 2>&1; echo "ab exit $?" >> gpurun_out/p_ab.log
grep -E "^AB|exit|Error" gpurun_out/p_ab.log | cut -c1-330
