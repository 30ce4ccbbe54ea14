#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/mg_check.py > gpurun_out/p_mg_check.log 2>&1; echo "exit $?" >> gpurun_out/p_mg_check.log
grep -E "MG_CHECK|exit|Error|identical" gpurun_out/p_mg_check.log | tail -5 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 40 --warmup 5 --strong-case none --no-cpu --fast-init 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3))"
