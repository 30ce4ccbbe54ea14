#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "whole", round(d["roofline"]["whole_step_frac"],3), "e2e", round(d["e2e"]["value"]), "fast", round(d["fast_mode"]["value"]), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("strong", {k:v for k,v in (d.get("strong") or {}).items() if k in ("ms_per_coarse_step","mlups_true","Cd","error")})
print("cpu", d.get("cpu_baseline",{}).get("value"))
PY
