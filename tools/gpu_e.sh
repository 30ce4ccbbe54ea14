#!/bin/bash
# 1-GPU pass: new tests (graphs, export, checkpoint shards, C ABI, shared-GPU two-process check), sphere Re=1M with and without
# graphs, fork_max_blocks on one GPU, T_1 of the strong-scaling case in both modes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -6 gpurun_out/e_pytest.log
for v in "graphs=0" "graphs=1"; do
  for m in strict fast; do
    timeout 300 python tools/run_case_mg.py sphere_re1m 200 --fp-mode $m --variant "$v" > gpurun_out/e_sphere_${m}_${v/=/}.log 2>&1
    grep RESULT gpurun_out/e_sphere_${m}_${v/=/}.log | cut -c1-200
  done
done
timeout 600 python tools/run_case_mg.py bunny 8 --fp-mode strict --uniform-start --variant "graphs=0" --variant "graphs=0,fork_max_blocks=1000000" --variant "graphs=1" --variant "graphs=1,fork_max_blocks=1000000" > gpurun_out/e_bunny_1gpu_variants.log 2>&1
grep RESULT gpurun_out/e_bunny_1gpu_variants.log | cut -c1-220
timeout 900 python tools/run_case_mg.py bunny_fine 6 --fp-mode strict --uniform-start --profile 2 --variant "graphs=0" --variant "fork_max_blocks=1000000" > gpurun_out/e_bunny_fine_1gpu_strict.log 2>&1
grep RESULT gpurun_out/e_bunny_fine_1gpu_strict.log | cut -c1-220
timeout 600 python tools/run_case_mg.py bunny_fine 6 --fp-mode fast --uniform-start > gpurun_out/e_bunny_fine_1gpu_fast.log 2>&1
grep RESULT gpurun_out/e_bunny_fine_1gpu_fast.log | cut -c1-220
