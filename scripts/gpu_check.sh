#!/bin/bash
# usage: bash scripts/gpu_check.sh TAG [pytest-args]   -- parity tests, default bench, per-kernel profile; outputs tagged gpurun_out/r2_*_TAG.*
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 "$@" > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_gpu_$TAG.log | tail -15
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; cut -c1-330 gpurun_out/r2_bench_$TAG.json; grep -v Warn gpurun_out/r2_bench_$TAG.err | tail -3
TOP=90 timeout 200 python scripts/kprof.py > gpurun_out/r2_kprof_$TAG.log 2>&1; grep -v -i warn gpurun_out/r2_kprof_$TAG.log | head -64
