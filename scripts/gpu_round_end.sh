#!/bin/bash
# What a round-end GPU call looks like (run under gpurun): parity tests, smoke, bench, launch list, one ncu full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_l3_fwd_tc3' -s 2 -c 1 -o gpurun_out/prof_l3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_l3.ncu-rep --page raw --csv > gpurun_out/prof_l3_raw.csv 2>/dev/null
