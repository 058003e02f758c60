#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.log; grep -o '"gpu_launches": [0-9]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*\|"traffic": [0-9]*' gpurun_out/bench.log
