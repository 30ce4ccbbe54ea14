#!/bin/bash
# 8-GPU pass: strong scaling of BASELINE config 5 (bunny at surface_resolution 1300, 6 levels, 339 M cells), partition rules A/B on the
# same processes (the domain is built once)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 1200 $TR tools/run_case_mg.py bunny_fine 6 --fp-mode strict --uniform-start --profile 2 --json gpurun_out/g_bunny_fine_8gpu.json \
  --variant "partition=rcb_yz" --variant "plan" --variant "partition=rcb" --variant "partition=rcb_yz,fork_max_blocks=40000" \
  --variant "partition=rcb_yz,fp=fast" --variant "plan,fp=fast" > gpurun_out/g_bunny_fine_8gpu.log 2>&1
echo "exit $?"
grep -E "RESULT|domain build" gpurun_out/g_bunny_fine_8gpu.log | cut -c1-260
