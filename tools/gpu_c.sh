#!/bin/bash
# third GPU pass: kernel-variant A/B (ticket-scheduled TMA, cta_threads) in both FP modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_k1_features_gpu.py tests/test_k1_single_level_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "variant or cta_threads" > gpurun_out/c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -3 gpurun_out/c_pytest.log
run() { # name, args...
  local name=$1; shift
  timeout 400 python bench.py --steps 40 --warmup 5 --strong-case none --no-cpu "$@" > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "classes", {k: round(x,3) for k,x in d["roofline"]["class_ms_per_step"].items()})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/c_bench_$name.err").read()[-1500:])
PY
}
run strict_tma --fp-mode strict --option strict_kernel=tma
run strict_128 --fp-mode strict --option cta_threads=128
run strict_64 --fp-mode strict --option cta_threads=64
run fast_tma --fp-mode fast --option fast_kernel=tma
run fast_128 --fp-mode fast --option cta_threads=128
run fast_64 --fp-mode fast --option cta_threads=64
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k1_strict -s 2 -c 1 -o gpurun_out/c_prof_strict_tma python bench.py --fast-init --steps 2 --warmup 1 --no-cpu --strong-case none --fp-mode strict --nb 32 --option strict_kernel=tma > gpurun_out/c_ncu_tma.log 2>&1
