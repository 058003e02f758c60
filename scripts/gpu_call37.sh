#!/bin/bash
mkdir -p gpurun_out
PGPD_L3_DEBUG=1 PGPD_L3_VERSION=3 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug_v3.log 2>&1; head -5 gpurun_out/l3_debug_v3.log
PGPD_L3_VERSION=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu_v3.log 2>&1; tail -2 gpurun_out/pytest_gpu_v3.log
PGPD_L3_VERSION=3 timeout 200 python scripts/kprof.py > gpurun_out/kprof_v3.log 2>&1; grep -v Warn gpurun_out/kprof_v3.log | head -5
