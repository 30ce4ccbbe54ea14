#!/bin/bash
mkdir -p gpurun_out
run() { # name, args...
  local name=$1; shift
  timeout 400 python bench.py --steps 40 --warmup 5 --strong-case none --no-cpu "$@" > gpurun_out/l_bench_$name.json 2> gpurun_out/l_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/l_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "classes", {k: round(x,3) for k,x in d["roofline"]["class_ms_per_step"].items()})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/l_bench_$name.err").read()[-1500:])
PY
}
run strict_occ4 --fp-mode strict --option strict_occupancy=4
run strict_occ5 --fp-mode strict --option strict_occupancy=5
run strict_occ4_pf512 --fp-mode strict --option strict_occupancy=4 --option prefetch_distance=512
run strict_occ5_pf512 --fp-mode strict --option strict_occupancy=5 --option prefetch_distance=512
run strict_occ4_pf2048 --fp-mode strict --option strict_occupancy=4 --option prefetch_distance=2048
run fast_pf0 --fp-mode fast
run fast_pf512 --fp-mode fast --option prefetch_distance=512
run fast_pf2048 --fp-mode fast --option prefetch_distance=2048
timeout 300 python -m pytest tests/test_k1_single_level_gpu.py tests/test_k1_features_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/l_pytest.log 2>&1; tail -2 gpurun_out/l_pytest.log
