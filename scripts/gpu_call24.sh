#!/bin/bash
mkdir -p gpurun_out
for w in fwd bwdb; do WHICH=$w timeout 200 python scripts/stream_debug.py; done > gpurun_out/stream_dbg.log 2>&1
WHICH=bwda PGPD_TC_MASK=0x1F timeout 200 python scripts/stream_debug.py >> gpurun_out/stream_dbg.log 2>&1
timeout 200 python scripts/kprof.py > gpurun_out/kprof_base.log 2>&1
for v in np8_eb16 np8_eb32 np4_eb16; do PGPD_LIB=build/variants/libpgpd_$v.so TOP=8 timeout 200 python scripts/kprof.py > gpurun_out/kprof_$v.log 2>&1; done
PGPD_LIB=build/variants/libpgpd_np8_eb16.so WHICH=bwdb timeout 200 python scripts/stream_debug.py > gpurun_out/stream_dbg_np8.log 2>&1
cat gpurun_out/stream_dbg.log; cat gpurun_out/kprof_base.log; head -12 gpurun_out/kprof_np*.log; cat gpurun_out/stream_dbg_np8.log
