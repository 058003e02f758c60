#!/bin/bash
mkdir -p gpurun_out
( for m in 0x3F 0x01 0x03 0x07 0x0B 0x1B 0x23 0x2B; do PGPD_TC_MASK=$m timeout 120 python scripts/diag_tc.py 48 1000; done ) > gpurun_out/diag.log 2>&1
( PGPD_TC_MASK=0x3F timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/diag_tc.py 6 300 ) > gpurun_out/racecheck.log 2>&1
( PGPD_TC_MASK=0x3F timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/diag_tc.py 6 300 ) > gpurun_out/memcheck.log 2>&1
cat gpurun_out/diag.log; tail -30 gpurun_out/racecheck.log; tail -30 gpurun_out/memcheck.log
