#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/kb_debug.py > gpurun_out/kb_debug.log 2>&1; cat gpurun_out/kb_debug.log | tail -8
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python scripts/kprof.py > gpurun_out/kprof.log 2>&1; grep -v Warn gpurun_out/kprof.log | head -34
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.log
