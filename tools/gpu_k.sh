#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
( time $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/k_bench_n$N.json 2> gpurun_out/k_bench_n$N.err ) 2> gpurun_out/k_bench_n$N.time
tail -3 gpurun_out/k_bench_n$N.time
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/k_bench_n$N.json") if l.startswith("{")][-1])
    s=d.pop("strong"); print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
    print("STRONG", {k:v for k,v in (s or {}).items() if k not in ("rank0_levels_ms","all_ranks")})
except Exception as e:
    print("failed", e); print(open("gpurun_out/k_bench_n$N.err").read()[-3000:])
PY
( time $TR bench.py --impl reference --gpus $N --steps 100 --warmup 10 > gpurun_out/k_bench_ref_n$N.json 2> gpurun_out/k_bench_ref_n$N.err ) 2> gpurun_out/k_bench_ref_n$N.time
tail -3 gpurun_out/k_bench_ref_n$N.time; cut -c1-200 gpurun_out/k_bench_ref_n$N.json
