#!/bin/bash
# first GPU call: tcgen05 probe, GPU parity tests, a first bench line, kernel launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
NV=$(./tests/tc_probe/tc_probe.bin)
for v in $(seq 0 $((NV-1))); do timeout 60 ./tests/tc_probe/tc_probe.bin $v; echo "exit=$?"; done > gpurun_out/probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_simt.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/probe.log gpurun_out/pytest_gpu.log gpurun_out/bench.log
