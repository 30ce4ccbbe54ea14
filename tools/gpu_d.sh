#!/bin/bash
# 2-GPU pass: one-process-per-GPU parity (NCCL plumbing, NVLink pulls, native barrier) in both FP modes and every partition rule,
# then the shipped bunny (79.5 M cells, 5 levels) on 1 and 2 GPUs with three partition rules
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR tools/mg_check.py > gpurun_out/d_mg_check.log 2>&1; echo "exit $?" >> gpurun_out/d_mg_check.log
grep -c "bit-identical=True" gpurun_out/d_mg_check.log; grep -c "bit-identical=False" gpurun_out/d_mg_check.log; grep "MG_CHECK\|aero" gpurun_out/d_mg_check.log | tail -5
timeout 600 python tools/run_case_mg.py bunny 6 --fp-mode strict --ramp 16 --profile 2 > gpurun_out/d_bunny_1gpu_strict.log 2>&1
grep RESULT gpurun_out/d_bunny_1gpu_strict.log
timeout 600 python tools/run_case_mg.py bunny 6 --fp-mode fast --ramp 16 > gpurun_out/d_bunny_1gpu_fast.log 2>&1
grep RESULT gpurun_out/d_bunny_1gpu_fast.log
timeout 900 $TR tools/run_case_mg.py bunny 6 --fp-mode strict --ramp 16 --profile 2 --variant plan --variant partition=rcb_yz --variant partition=rcb --variant "partition=rcb_yz,fork_max_blocks=200000" > gpurun_out/d_bunny_2gpu_strict.log 2>&1
grep "RESULT" gpurun_out/d_bunny_2gpu_strict.log
tail -3 gpurun_out/d_bunny_2gpu_strict.log | cut -c1-400
