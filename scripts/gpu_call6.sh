#!/bin/bash
mkdir -p gpurun_out
( for m in 0x3F; do PGPD_TC_MASK=$m timeout 120 python scripts/diag_tc.py 48 1000; PGPD_TC_MASK=$m timeout 120 python scripts/diag_tc.py 7 333; done ) > gpurun_out/diag.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/diag.log; tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log
