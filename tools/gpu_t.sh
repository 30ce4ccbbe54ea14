#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 > gpurun_out/t_bench_n2.json 2> gpurun_out/t_bench_n2.err ) 2> gpurun_out/t_bench_n2.time; echo "bench exit $?"
tail -3 gpurun_out/t_bench_n2.time
python - <<'PY'
import json
d=json.loads(open("gpurun_out/t_bench_n2.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), "clocks", d["clocks"])
print("strong", {k:v for k,v in (d.get("strong") or {}).items() if k in ("ms_per_coarse_step","mlups_true","Cd","Cl","error","n_gpus","efficiency_vs_committed_T1")})
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 8 --warmup 1 > gpurun_out/t_bench_ref_n2.json 2> gpurun_out/t_bench_ref_n2.err ) 2> gpurun_out/t_bench_ref_n2.time; tail -3 gpurun_out/t_bench_ref_n2.time; cut -c1-400 gpurun_out/t_bench_ref_n2.json
