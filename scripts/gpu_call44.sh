#!/bin/bash
mkdir -p gpurun_out
B=16 N=1024 PGPD_L3_SPLIT=1 timeout 120 python scripts/kb_check.py gpurun_out/g_s1.npz 2>&1 | tail -1
B=16 N=1024 PGPD_L3_SPLIT=0 timeout 120 python scripts/kb_check.py gpurun_out/g_s0.npz 2>&1 | tail -1
python scripts/kb_cmp.py gpurun_out/g_s1.npz gpurun_out/g_s0.npz | grep -E "worst|logp|conv3.weight|stn.conv1.weight"
PGPD_L3_DEBUG=1 PGPD_L3_SPLIT=1 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug_v3s.log 2>&1; head -5 gpurun_out/l3_debug_v3s.log
PGPD_L3_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu_split.log 2>&1; tail -1 gpurun_out/pytest_gpu_split.log
PGPD_L3_SPLIT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' | head -4
rm -f gpurun_out/*.npz
