#!/bin/bash
mkdir -p gpurun_out
PGPD_L3_DEBUG=1 PGPD_L3_VERSION=3 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug_v3.log 2>&1; head -5 gpurun_out/l3_debug_v3.log
PGPD_L3_DEBUG=1 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug_v1.log 2>&1; head -5 gpurun_out/l3_debug_v1.log
PGPD_L3_VERSION=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu_v3.log 2>&1; tail -1 gpurun_out/pytest_gpu_v3.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu_v1.log 2>&1; tail -1 gpurun_out/pytest_gpu_v1.log
