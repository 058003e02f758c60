#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for v in nopdl pdl; do PGPD_LIB=build/variants/libpgpd_$v.so timeout 120 python scripts/step_time.py 2>&1 | grep "graph step" | tee -a gpurun_out/r2_ab_pdl.log; done; done
bash scripts/gpu_check.sh c7
