#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/run_case_mg.py bunny 6 --fp-mode strict --uniform-start --profile 2 --variant "serial_prepass=1,prepass=thread,graphs=0" --variant "serial_prepass=1,prepass=block,graphs=0" --variant "prepass=thread" --variant "prepass=block" > gpurun_out/h_bunny_prepass.log 2>&1
grep -E "RESULT|rank 0" gpurun_out/h_bunny_prepass.log | cut -c1-330 | sed 's/per level \[ms\/coarse step\]: L1.*| L5/L5/'
timeout 400 python bench.py --steps 40 --warmup 5 --strong-case none --no-cpu --fp-mode strict > gpurun_out/h_bench_strict.json 2> gpurun_out/h_bench_strict.err
timeout 400 python bench.py --steps 40 --warmup 5 --strong-case none --no-cpu --fp-mode fast > gpurun_out/h_bench_fast.json 2> gpurun_out/h_bench_fast.err
for f in h_bench_strict h_bench_fast; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "step frac", round(d["roofline"]["whole_step_frac"],3), "classes", {k: round(v,3) for k,v in d["roofline"]["class_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/$f.err").read()[-800:])
PY
done
timeout 300 python -m pytest tests/test_k1_single_level_gpu.py tests/test_k1_features_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1; tail -2 gpurun_out/h_pytest.log
