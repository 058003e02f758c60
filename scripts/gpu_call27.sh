#!/bin/bash
mkdir -p gpurun_out
for cfg in "24 1000" "5 333" "16 1024"; do set -- $cfg
  B=$1 N=$2 PGPD_KA=1 timeout 120 python scripts/kb_check.py gpurun_out/g_new_$1_$2.npz 2>&1 | tail -3
  B=$1 N=$2 PGPD_KA=0 timeout 120 python scripts/kb_check.py gpurun_out/g_old_$1_$2.npz 2>&1 | tail -3
  python scripts/kb_cmp.py gpurun_out/g_new_$1_$2.npz gpurun_out/g_old_$1_$2.npz
done > gpurun_out/ka_check.log 2>&1
cat gpurun_out/ka_check.log
timeout 200 python scripts/kprof.py > gpurun_out/kprof_ka.log 2>&1; grep -v Warn gpurun_out/kprof_ka.log | head -34
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_kb_tc|k_ka_tc' -s 2 -c 2 -o gpurun_out/prof_kab -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_kab.log 2>&1
tail -2 gpurun_out/ncu_kab.log | cut -c1-200
rm -f gpurun_out/*.npz
