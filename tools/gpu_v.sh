#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 tools/run_case_mg.py bunny_fine 6 --fp-mode strict --uniform-start \
   --variant "partition=rcb_yz" --variant "partition=rcb_yz,fp=fast" --variant "partition=rcb_yz" --profile 2 --json gpurun_out/v_bunny_fine_${N}gpu.json > gpurun_out/v_bunny_fine_${N}gpu.log 2>&1; echo "exit $?" >> gpurun_out/v_bunny_fine_${N}gpu.log
grep -E "RESULT|exit|Error" gpurun_out/v_bunny_fine_${N}gpu.log | cut -c1-300
