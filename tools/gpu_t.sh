#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N > gpurun_out/t_bench_n$N.json 2> gpurun_out/t_bench_n$N.err ) 2> gpurun_out/t_bench_n$N.time; echo "bench exit $?"
tail -3 gpurun_out/t_bench_n$N.time
python - <<PY
import json
d=json.loads(open("gpurun_out/t_bench_n$N.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), "clocks", d["clocks"])
print("strong", {k:v for k,v in (d.get("strong") or {}).items() if k in ("ms_per_coarse_step","mlups_true","Cd","Cl","error","n_gpus","efficiency_vs_committed_T1")})
PY
