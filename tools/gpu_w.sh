#!/bin/bash
# round-2 final evidence: ncu launch list of the bench command, DRAM traffic of the dominant kernels at 512^3, one --set full capture
mkdir -p gpurun_out
timeout 600 python bench.py --fast-init --steps 4 --warmup 3 --strong-case none --no-cpu > gpurun_out/w_bench_fastinit.json 2> gpurun_out/w_bench_fastinit.err; echo "plain run exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w_launches_bench_512cube.csv python bench.py --fast-init --steps 4 --warmup 3 --strong-case none --no-cpu > gpurun_out/w_ncu1.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"k1_" -c 24 --csv --log-file gpurun_out/w_dram_traffic_k1_512cube.csv \
  python tools/ab_box.py --nb 64 --steps 2 --warmup 2 --repeat 1 --profile-steps 0 --e2e-steps 0 "default|strict|" "default|fast|" "unmerged|strict|merge_face=0" > gpurun_out/w_ncu2.log 2>&1; echo "traffic exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_strict_mixed" -s 4 -c 1 -o gpurun_out/w_k1_strict_mixed_full -f python tools/ab_box.py --nb 32 --steps 4 --warmup 4 --repeat 1 --profile-steps 0 --e2e-steps 0 "default|strict|" > gpurun_out/w_ncu3.log 2>&1; echo "full exit $?"
ncu -i gpurun_out/w_k1_strict_mixed_full.ncu-rep --page raw --csv > gpurun_out/w_ncu_full_k1_strict_mixed_256cube_raw.csv 2>/dev/null
ls -la gpurun_out/w_* | cut -c1-150
