#!/bin/bash
mkdir -p gpurun_out
for cfg in "24 1000" "5 333" "16 1024"; do set -- $cfg
  B=$1 N=$2 PGPD_TC_MASK=0x3F timeout 120 python scripts/kb_check.py gpurun_out/g_tc_$1_$2.npz 2>&1 | tail -3
  B=$1 N=$2 PGPD_TC_MASK=0x2F timeout 120 python scripts/kb_check.py gpurun_out/g_ref_$1_$2.npz 2>&1 | tail -3
  python scripts/kb_cmp.py gpurun_out/g_tc_$1_$2.npz gpurun_out/g_ref_$1_$2.npz
done > gpurun_out/kb_check.log 2>&1
cat gpurun_out/kb_check.log
timeout 200 python scripts/kprof.py > gpurun_out/kprof_kb.log 2>&1; grep -v Warn gpurun_out/kprof_kb.log | head -40
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/*.npz
