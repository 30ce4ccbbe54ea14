#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/run_case_mg.py bunny 6 --fp-mode strict --uniform-start --profile 2 --variant "verbose=0" --variant "p.wall_model_active=0" --variant "block_order=xslab12" --variant "fp=fast" --variant "fp=fast,block_order=xslab12" > gpurun_out/r_bunny.log 2>&1; echo "exit $?" >> gpurun_out/r_bunny.log
grep -E "RESULT|rank 0|exit|Error" gpurun_out/r_bunny.log | cut -c1-900
