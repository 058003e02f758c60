#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_l3_fwd_tc3|k_kf_tc' -s 4 -c 2 -o gpurun_out/prof_l3kf -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log | cut -c1-160
ncu -i gpurun_out/prof_l3kf.ncu-rep --page raw --csv > gpurun_out/prof_l3kf_raw.csv 2>/dev/null
timeout 600 python scripts/infer_sweep.py > gpurun_out/infer_sweep.log 2>&1; cut -c1-330 gpurun_out/infer_sweep.log | head -8
rm -f gpurun_out/prof_l3kf.ncu-rep gpurun_out/prof_main.ncu-rep
