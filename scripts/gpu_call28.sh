#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/kb_debug.py > gpurun_out/kb_debug.log 2>&1; cat gpurun_out/kb_debug.log | tail -9
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python scripts/kprof.py > gpurun_out/kprof_kf.log 2>&1; grep -v Warn gpurun_out/kprof_kf.log | head -16
