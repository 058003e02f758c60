#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
PGPD_L3_DEBUG=1 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/l3_debug.log; grep -E "^E  |passed|failed" gpurun_out/pytest_gpu.log | head; cut -c1-250 gpurun_out/bench.log; grep -o '"kernel_ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench.log
