#!/bin/bash
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'k_stream_tc|k_accum_tc' -s 10 -c 5 -o gpurun_out/prof_stream -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log | cut -c1-300
ls -la gpurun_out/
