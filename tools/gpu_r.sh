#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_options.py "feature_first=0" > gpurun_out/r_check.log 2>&1; grep -E "CHECK|Error" gpurun_out/r_check.log | cut -c1-200
timeout 1200 python tools/run_case_mg.py bunny 8 --fp-mode strict --uniform-start --variant "feature_first=1" --variant "feature_first=0" --variant "feature_first=1" --variant "feature_first=0" --variant "feature_first=1,fp=fast" --variant "feature_first=0,fp=fast" > gpurun_out/r_bunny2.log 2>&1; echo "exit $?" >> gpurun_out/r_bunny2.log
grep -E "RESULT|exit|Error" gpurun_out/r_bunny2.log | cut -c1-200
timeout 1200 python tools/run_case_mg.py wing5 8 --fp-mode strict --uniform-start --variant "feature_first=1" --variant "feature_first=0"  --variant "feature_first=1" --variant "feature_first=0" > gpurun_out/r_wing2.log 2>&1; echo "exit $?" >> gpurun_out/r_wing2.log
grep -E "RESULT|exit|Error" gpurun_out/r_wing2.log | cut -c1-200
