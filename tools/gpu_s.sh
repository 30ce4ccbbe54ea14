#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -4 gpurun_out/s_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/s_bench_ref.json 2> gpurun_out/s_bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "whole", round(d["roofline"]["whole_step_frac"],3), "e2e", round(d["e2e"]["value"]), "fast", round(d["fast_mode"]["value"]), "clocks", d["clocks"])
print("strong", {k:v for k,v in (d.get("strong") or {}).items() if k in ("ms_per_coarse_step","mlups_true","Cd","Cl","error","n_gpus")})
print("cpu", d.get("cpu_baseline"))
r=json.loads(open("gpurun_out/s_bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
python -c "import __graft_entry__ as g; g.smoke()"
