#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/kb_debug.py > gpurun_out/kb_debug.log 2>&1; tail -9 gpurun_out/kb_debug.log
WHICH=ka PGPD_TC_MASK=0x2F timeout 200 python scripts/kb_debug.py > gpurun_out/ka_debug.log 2>&1; tail -9 gpurun_out/ka_debug.log
for cfg in "24 1000" "5 333"; do set -- $cfg
  B=$1 N=$2 PGPD_TC_MASK=0x3F timeout 120 python scripts/kb_check.py gpurun_out/g_tc_$1_$2.npz 2>&1 | tail -1
  B=$1 N=$2 PGPD_TC_MASK=0x2F timeout 120 python scripts/kb_check.py gpurun_out/g_ref_$1_$2.npz 2>&1 | tail -1
  python scripts/kb_cmp.py gpurun_out/g_tc_$1_$2.npz gpurun_out/g_ref_$1_$2.npz | tail -12
done > gpurun_out/kb_check.log 2>&1
grep -E "worst|conv1.weight|conv2.weight|bn1.weight" gpurun_out/kb_check.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python scripts/kprof.py > gpurun_out/kprof.log 2>&1; grep -v Warn gpurun_out/kprof.log | head -12
rm -f gpurun_out/*.npz
