#!/bin/bash
# layer-3 kernel diagnosis on ONE box: cycle accounting of debug builds (build/variants/libpgpd_dbg*.so), then interleaved timing of
# the matching product builds.  usage: bash scripts/gpu_l3diag.sh "dbg dbgnorot" "base norot"
mkdir -p gpurun_out
[ -f scripts/l3diag.conf ] && source scripts/l3diag.conf && set -- "$D1" "$D2" "$D3" "$D4"
for v in $1; do
  echo "=== counters: $v" | tee -a gpurun_out/r2_l3diag.log
  PGPD_L3_DEBUG=1 PGPD_LIB=build/variants/libpgpd_$v.so timeout 120 python scripts/l3_debug.py 2>&1 | grep -v -i warn | tee -a gpurun_out/r2_l3diag.log
done
PAT="k_l3_fwd|eager fwd" bash scripts/gpu_ab.sh $2
if [ -n "$3" ]; then
  echo "=== ka / kb counters: $3" | tee -a gpurun_out/r2_l3diag.log
  PGPD_LIB=build/variants/libpgpd_$3.so timeout 120 python scripts/kakb_debug.py 2>&1 | grep -v -i warn | tee -a gpurun_out/r2_l3diag.log
fi
if [ -n "$4" ]; then timeout 300 python -m pytest tests/test_dual.py -m gpu -q 2>&1 | tail -3; fi
