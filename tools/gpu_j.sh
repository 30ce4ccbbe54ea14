#!/bin/bash
mkdir -p gpurun_out
( time python bench.py > gpurun_out/j_bench_default.json 2> gpurun_out/j_bench_default.err ) 2> gpurun_out/j_bench_default.time
tail -3 gpurun_out/j_bench_default.time
python - <<PY
import json
d=json.loads(open("gpurun_out/j_bench_default.json").read().strip().splitlines()[-1])
s=d.pop("strong"); print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk not in ("what","workload","parity","multi_gpu","traffic_source","peak_source","sample")}) for k,v in d.items()})
print("STRONG", {k:v for k,v in (s or {}).items() if k not in ("rank0_levels_ms","all_ranks")})
PY
( time python bench.py --impl reference > gpurun_out/j_bench_reference.json 2> gpurun_out/j_bench_reference.err ) 2> gpurun_out/j_bench_reference.time
tail -3 gpurun_out/j_bench_reference.time; cut -c1-400 gpurun_out/j_bench_reference.json
timeout 300 python -m pytest tests/test_k1_features_gpu.py tests/test_cases_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/j_pytest.log 2>&1; tail -2 gpurun_out/j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
