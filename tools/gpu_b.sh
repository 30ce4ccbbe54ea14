#!/bin/bash
# second GPU pass: strict-kernel variants (register / stash / TMA): bit-identity tests, then bench of each, then ncu of the plain strict kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_k1_features_gpu.py tests/test_k1_single_level_gpu.py tests/test_virtual_ranks_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "variants or virtual_ranks_two_level" > gpurun_out/b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -4 gpurun_out/b_pytest.log
for v in reg stash tma; do
  timeout 400 python bench.py --steps 40 --warmup 5 --strong-case none --fp-mode strict --no-cpu --option strict_kernel=$v > gpurun_out/b_bench_strict_$v.json 2> gpurun_out/b_bench_strict_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_bench_strict_$v.json").read().strip().splitlines()[-1])
    print("strict $v", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "classes", {k: round(x,3) for k,x in d["roofline"]["class_ms_per_step"].items()})
except Exception as e:
    print("strict $v failed", e); print(open("gpurun_out/b_bench_strict_$v.err").read()[-1500:])
PY
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k1_strict -s 2 -c 1 -o gpurun_out/b_prof_strict_reg python bench.py --fast-init --steps 2 --warmup 1 --no-cpu --strong-case none --fp-mode strict --nb 32 > gpurun_out/b_ncu_reg.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k1_strict -s 2 -c 1 -o gpurun_out/b_prof_strict_tma python bench.py --fast-init --steps 2 --warmup 1 --no-cpu --strong-case none --fp-mode strict --nb 32 --option strict_kernel=tma > gpurun_out/b_ncu_tma.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
