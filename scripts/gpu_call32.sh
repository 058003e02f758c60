#!/bin/bash
mkdir -p gpurun_out
# A/B: heads on tensor cores vs CUDA cores (bit 6), same everything else
for cfg in "24 1000" "130 256"; do set -- $cfg
  B=$1 N=$2 PGPD_TC_MASK=0x7F timeout 120 python scripts/kb_check.py gpurun_out/g_tc_$1_$2.npz 2>&1 | tail -1
  B=$1 N=$2 PGPD_TC_MASK=0x3F timeout 120 python scripts/kb_check.py gpurun_out/g_ref_$1_$2.npz 2>&1 | tail -1
  python scripts/kb_cmp.py gpurun_out/g_tc_$1_$2.npz gpurun_out/g_ref_$1_$2.npz | grep -E "worst|logp|fc1.weight|fc2.weight|stn.conv3.weight"
done > gpurun_out/head_check.log 2>&1
cat gpurun_out/head_check.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python scripts/kprof.py > gpurun_out/kprof.log 2>&1; grep -v Warn gpurun_out/kprof.log | head -30
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.log
rm -f gpurun_out/*.npz
