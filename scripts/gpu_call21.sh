#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/infer_sweep.py > gpurun_out/infer_sweep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_l3_fwd_tc' -s 4 -c 2 -o gpurun_out/prof_l3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_l3.ncu-rep --page raw --csv > gpurun_out/prof_l3_raw.csv 2>/dev/null
cat gpurun_out/infer_sweep.log | cut -c1-400; tail -2 gpurun_out/ncu_full.log | cut -c1-200
