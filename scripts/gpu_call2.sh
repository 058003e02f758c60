#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/r2_pytest_gpu_c2.log 2>&1; tail -12 gpurun_out/r2_pytest_gpu_c2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; cut -c1-300 gpurun_out/r2_bench_c2.json; tail -3 gpurun_out/r2_bench_c2.err
TOP=90 timeout 200 python scripts/kprof.py > gpurun_out/r2_kprof_c2.log 2>&1; grep -v Warn gpurun_out/r2_kprof_c2.log | head -60
