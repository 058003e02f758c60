#!/bin/bash
mkdir -p gpurun_out
for cfg in "24 1000" "5 333" "16 1024"; do set -- $cfg
  B=$1 N=$2 PGPD_L3_VERSION=3 timeout 120 python scripts/kb_check.py gpurun_out/g_v3_$1_$2.npz 2>&1 | tail -2
  B=$1 N=$2 PGPD_L3_VERSION=1 timeout 120 python scripts/kb_check.py gpurun_out/g_v1_$1_$2.npz 2>&1 | tail -1
  python scripts/kb_cmp.py gpurun_out/g_v3_$1_$2.npz gpurun_out/g_v1_$1_$2.npz | grep -E "worst|logp|conv3.weight|bn3.weight|conv1.weight"
done > gpurun_out/l3v3_check.log 2>&1
cat gpurun_out/l3v3_check.log
PGPD_L3_VERSION=3 timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu_v3.log 2>&1; tail -4 gpurun_out/pytest_gpu_v3.log
PGPD_L3_VERSION=3 timeout 200 python scripts/kprof.py > gpurun_out/kprof_v3.log 2>&1; grep -v Warn gpurun_out/kprof_v3.log | head -8
timeout 200 python scripts/kprof.py > gpurun_out/kprof_v1.log 2>&1; grep -v Warn gpurun_out/kprof_v1.log | head -6
rm -f gpurun_out/*.npz
