#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_l3_fwd_tc|k_stream_tc|k_accum_tc' -s 12 -c 12 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
tail -5 gpurun_out/ncu_full.log
