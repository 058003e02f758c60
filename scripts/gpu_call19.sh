#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/bench.log
grep -E "^E  |passed|failed" gpurun_out/pytest_gpu.log | head; cut -c1-250 gpurun_out/bench.log; grep -o '"cpu_baseline.*' gpurun_out/bench.log | cut -c1-300
