#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python scripts/diag_tc.py 7 333; echo "exit=$?"; timeout 120 python scripts/diag_tc.py 48 1000; echo "exit=$?" ) > gpurun_out/diag.log 2>&1
if grep -q "exit=0" gpurun_out/diag.log && ! grep -q "<<<<<<" gpurun_out/diag.log; then echo "L3 v2 OK" >> gpurun_out/diag.log; else echo "L3 v2 BROKEN -> using v1" >> gpurun_out/diag.log; export PGPD_L3_VERSION=1; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
PGPD_L3_DEBUG=1 timeout 300 python scripts/l3_debug.py > gpurun_out/l3_debug.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/bench.log
PGPD_L3_VERSION=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v1.log 2>> gpurun_out/bench.err
timeout 300 python scripts/graph_try.py > gpurun_out/graph.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -4 gpurun_out/diag.log; cat gpurun_out/l3_debug.log; grep -E "^E  |passed|failed" gpurun_out/pytest_gpu.log | head; cut -c1-200 gpurun_out/bench.log; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench.log gpurun_out/bench_v1.log; tail -3 gpurun_out/graph.log
