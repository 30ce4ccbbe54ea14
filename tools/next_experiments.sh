#!/bin/bash
# The measurements queued at the end of round 1, as the exact commands.  Each block is ONE gpurun
# call; run them one at a time and copy what should be judged from gpurun_out/ into profiles/.  Not run by any test.
set -e
case "$1" in
1gpu)   # parity + bench + the two single-GPU experiments + DRAM traffic of K1 at 512^3   (~8 GPU-minutes)
  gpurun --timeout 900 -- 'python -m pytest tests -m gpu -x -q 2>&1 | tail -4;
    python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err;
    LUDWIG_FORK_FULL=1 python bench.py --no-cpu > gpurun_out/bench_fork_full.json 2> gpurun_out/bench_fork_full.err;
    timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k1_fast_kernel -c 4 --csv --log-file gpurun_out/k1_traffic_512cube.csv python bench.py --fast-init --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_traffic.log 2>&1;
    grep -h "\"value\"" gpurun_out/bench_default.json gpurun_out/bench_fork_full.json | cut -c1-160' ;;
emulate)  # partition-only bound of three partitions on ONE GPU   (~4 GPU-minutes each)
  gpurun --timeout 1500 -- 'python tools/emulate_ranks.py bunny_fine 8 2 > gpurun_out/emu_plan.log 2>&1; tail -1 gpurun_out/emu_plan.log;
    LUDWIG_PARTITION=rcb python tools/emulate_ranks.py bunny_fine 8 2 > gpurun_out/emu_rcb.log 2>&1; tail -1 gpurun_out/emu_rcb.log;
    python tools/emulate_ranks.py bunny_fine 8 2 --no-plan > gpurun_out/emu_noplan.log 2>&1; tail -1 gpurun_out/emu_noplan.log' ;;
2gpu)   # bit-identity of every multi-GPU variant (native / NCCL barrier, packed mirrors, plan, RCB)   (~2 GPU-minutes)
  gpurun --gpus 2 --timeout 300 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/mg_check.py > gpurun_out/mg2.log 2>&1; grep -c "bit-identical=True" gpurun_out/mg2.log; grep MG_CHECK gpurun_out/mg2.log' ;;
8gpu)   # strong scaling of config 5, one variant per call: $2 = extra environment, e.g. "LUDWIG_PARTITION=rcb"   (~9 GPU-minutes)
  gpurun --gpus 8 --timeout 420 -- "$2 LUDWIG_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/run_case_mg.py bunny_fine 4 > gpurun_out/bf8_\$(echo $2 | tr -c 'A-Za-z0-9\n' _).log 2>&1; grep -h 'RESULT\|^rank' gpurun_out/bf8_*.log | tail -9" ;;
*) echo "usage: $0 1gpu | emulate | 2gpu | 8gpu '<ENV=VALUE ...>'   (variants: LUDWIG_PARTITION=rcb, LUDWIG_NO_PLAN=1, LUDWIG_FORK_MAX_BLOCKS=100000, LUDWIG_REMOTE_ORDER=interleave, LUDWIG_HALO_MIRROR=1)" ;;
esac
