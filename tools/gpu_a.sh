#!/bin/bash
# first GPU pass of round 2: parity suite, bench in both FP modes, fork_full A/B, DRAM traffic of K1 at 512^3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/a_gpu.txt
(nproc; free -g | head -2) >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 --strong-case none --fp-mode fast > gpurun_out/a_bench_fast.json 2> gpurun_out/a_bench_fast.err
timeout 600 python bench.py --steps 50 --warmup 5 --strong-case none --fp-mode strict --no-cpu > gpurun_out/a_bench_strict.json 2> gpurun_out/a_bench_strict.err
timeout 600 python bench.py --steps 50 --warmup 5 --strong-case none --fp-mode fast --no-cpu --option fork_full=1 > gpurun_out/a_bench_fast_forkfull.json 2> gpurun_out/a_bench_fast_forkfull.err
for f in a_bench_fast a_bench_strict a_bench_fast_forkfull; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "step frac", round(d["roofline"]["whole_step_frac"],3), "classes", {k: round(v,3) for k,v in d["roofline"]["class_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]), "other", d.get("strict_mode") or d.get("fast_mode"))
except Exception as e:
    print("$f failed", e)
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k1_ -c 6 --csv --log-file gpurun_out/a_k1_traffic_512cube_fast.csv python bench.py --fast-init --steps 2 --warmup 1 --no-cpu --strong-case none --fp-mode fast > gpurun_out/a_ncu_fast.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k1_ -c 6 --csv --log-file gpurun_out/a_k1_traffic_512cube_strict.csv python bench.py --fast-init --steps 2 --warmup 1 --no-cpu --strong-case none --fp-mode strict > gpurun_out/a_ncu_strict.log 2>&1
tail -4 gpurun_out/a_k1_traffic_512cube_fast.csv | cut -c1-300
