#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k1_|ghost_interp|bouzidi" -s 400 -c 160 --csv --log-file gpurun_out/o_launches_bunny_strict.csv python tools/run_case_mg.py bunny 2 --fp-mode strict --uniform-start --variant "graphs=0,single_stream=1" > gpurun_out/o_ncu.log 2>&1
tail -3 gpurun_out/o_ncu.log | cut -c1-200
wc -l gpurun_out/o_launches_bunny_strict.csv
