#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_options.py "merge_face=0" > gpurun_out/p_check.log 2>&1; echo "check exit $?" >> gpurun_out/p_check.log
grep -E "CHECK|exit|Error" gpurun_out/p_check.log | cut -c1-300
timeout 900 python tools/ab_box.py --nb 64 --steps 40 --repeat 3 "merge1|strict|merge_face=1" "merge0|strict|merge_face=0"  > gpurun_out/p_ab.log 2>&1; echo "ab exit $?" >> gpurun_out/p_ab.log
grep -E "^AB|exit|Error" gpurun_out/p_ab.log | cut -c1-330
