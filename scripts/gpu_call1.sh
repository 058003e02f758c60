#!/bin/bash
# round-2 call 1: new parity tests, drop-in proof, sanitizers, bench at every BASELINE config, per-kernel profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2_pytest_gpu_c1.log 2>&1; tail -15 gpurun_out/r2_pytest_gpu_c1.log
bash scripts/dropin_gpu.sh
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; cut -c1-400 gpurun_out/r2_bench_c1.json
timeout 200 python bench.py --steps 20 --warmup 5 --classes 3 --no-cpu-baseline > gpurun_out/r2_bench_c1_k3.json 2>> gpurun_out/r2_bench_c1.err; cut -c1-200 gpurun_out/r2_bench_c1_k3.json
timeout 200 python bench.py --steps 20 --warmup 5 --batch 128 --points 2048 --no-cpu-baseline > gpurun_out/r2_bench_c1_128x2048.json 2>> gpurun_out/r2_bench_c1.err; cut -c1-200 gpurun_out/r2_bench_c1_128x2048.json
timeout 300 python bench.py --steps 20 --warmup 5 --config infer > gpurun_out/r2_bench_c1_infer.json 2>> gpurun_out/r2_bench_c1.err; cut -c1-300 gpurun_out/r2_bench_c1_infer.json
timeout 300 python bench.py --steps 20 --warmup 5 --config tower > gpurun_out/r2_bench_c1_tower.json 2>> gpurun_out/r2_bench_c1.err; cut -c1-300 gpurun_out/r2_bench_c1_tower.json
tail -5 gpurun_out/r2_bench_c1.err
TOP=90 timeout 200 python scripts/kprof.py > gpurun_out/r2_kprof_c1.log 2>&1; tail -3 gpurun_out/r2_kprof_c1.log
bash scripts/sanitize_gpu.sh
