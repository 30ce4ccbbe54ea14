#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/run_case_mg.py bunny 8 --fp-mode strict --uniform-start --variant "verbose=0" --variant "fork_max_blocks=0" --variant "fork_max_blocks=40000"  --variant "fork_max_blocks=40000,serial_prepass=1" --variant "verbose=0" --variant "fork_max_blocks=0" --variant "fork_max_blocks=40000" --variant "fork_max_blocks=40000,serial_prepass=1" > gpurun_out/r_bunny2.log 2>&1; echo "exit $?" >> gpurun_out/r_bunny2.log
grep -E "RESULT|exit|Error" gpurun_out/r_bunny2.log | cut -c1-200
timeout 1200 python tools/run_case_mg.py wing5 8 --fp-mode strict --uniform-start --variant "verbose=0" --variant "fork_max_blocks=0" --variant "fork_max_blocks=40000"  --variant "verbose=0" --variant "fork_max_blocks=0" --variant "fork_max_blocks=40000" > gpurun_out/r_wing2.log 2>&1; echo "exit $?" >> gpurun_out/r_wing2.log
grep -E "RESULT|exit|Error" gpurun_out/r_wing2.log | cut -c1-200
