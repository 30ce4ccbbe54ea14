#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/run_case_mg.py bunny 8 --fp-mode strict --uniform-start \
   --variant "partition=rcb_yz" --variant "partition=rcb_yz,block_order=morton" --variant "partition=rcb_yz,block_order=xslab8" --variant "partition=rcb_yz,fp=fast" --variant "partition=rcb_yz,fp=fast,block_order=morton" > gpurun_out/t_bunny_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/t_bunny_2gpu.log
grep -E "RESULT|exit|Error" gpurun_out/t_bunny_2gpu.log | cut -c1-400
