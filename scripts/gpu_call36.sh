#!/bin/bash
mkdir -p gpurun_out
PGPD_L3_DEBUG=1 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug_v1.log 2>&1; tail -10 gpurun_out/l3_debug_v1.log
PGPD_L3_DEBUG=1 PGPD_L3_VERSION=3 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug_v3.log 2>&1; tail -10 gpurun_out/l3_debug_v3.log
