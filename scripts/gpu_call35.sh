#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./build/mma_time_pair.bin > gpurun_out/mma_time_pair.log 2>&1; cat gpurun_out/mma_time_pair.log
# clocks while the L3 kernel dominates: run the eval forward (L3-heavy) in a loop and sample
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 100 > gpurun_out/clocks_bench.csv &
SMI=$!
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench200.log 2> gpurun_out/bench200.err
kill $SMI
cut -c1-200 gpurun_out/bench200.log; sort gpurun_out/clocks_bench.csv | uniq -c | sort -rn | head -8
PGPD_L3_VERSION=3 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v3.log 2> gpurun_out/bench_v3.err; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/bench_v3.log | head -3
