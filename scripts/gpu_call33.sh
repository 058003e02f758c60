#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
PGPD_L3_DEBUG=1 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug.log 2>&1; tail -12 gpurun_out/l3_debug.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.log; grep -o '"roofline".*' gpurun_out/bench.log | cut -c1-400
