#!/bin/bash
# A/B timing of kernel variants built into build/variants/*.so on ONE box: scripts/kprof.py with PGPD_LIB, interleaved, twice
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    PGPD_LIB=build/variants/libpgpd_$v.so TOP=14 timeout 120 python scripts/kprof.py 2>&1 | grep -E "${PAT:-eager fwd|k_l3_fwd}" | sed "s/^/[$v $rep] /" | tee -a gpurun_out/r2_ab2.log
  done
done
