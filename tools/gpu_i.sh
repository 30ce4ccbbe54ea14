#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_k1_single_level_gpu.py tests/test_k1_features_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/i_pytest.log 2>&1; tail -2 gpurun_out/i_pytest.log
for occ in 4 5 6; do
  timeout 400 python bench.py --steps 40 --warmup 5 --strong-case none --no-cpu --fp-mode strict --option strict_occupancy=$occ > gpurun_out/i_bench_strict_occ$occ.json 2> gpurun_out/i_bench_strict_occ$occ.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/i_bench_strict_occ$occ.json").read().strip().splitlines()[-1])
    print("strict occ $occ", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "step frac", round(d["roofline"]["whole_step_frac"],3), "classes", {k: round(v,3) for k,v in d["roofline"]["class_ms_per_step"].items()})
except Exception as e:
    print("occ $occ failed", e); print(open("gpurun_out/i_bench_strict_occ$occ.err").read()[-800:])
PY
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k1_strict -s 2 -c 1 -o gpurun_out/i_prof_strict_occ5 python bench.py --fast-init --steps 2 --warmup 1 --no-cpu --strong-case none --fp-mode strict --nb 32 --option strict_occupancy=5 > gpurun_out/i_ncu.log 2>&1
timeout 600 python tools/run_case_mg.py bunny 6 --fp-mode strict --uniform-start --variant "strict_occupancy=4" --variant "strict_occupancy=5" --variant "strict_occupancy=6" > gpurun_out/i_bunny_occ.log 2>&1
grep RESULT gpurun_out/i_bunny_occ.log | cut -c1-150
