#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "eval or inference or shipped or simt_flag or large_magnitude or nan or reference_error" > gpurun_out/r2_pytest_fused.log 2>&1; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_fused.log | tail -15
timeout 300 python bench.py --steps 20 --warmup 5 --config tower --no-cpu-baseline > gpurun_out/r2_bench_fused_tower.json 2> gpurun_out/r2_bench_fused.err; cut -c1-250 gpurun_out/r2_bench_fused_tower.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fused_tower.json')); print(d['ms_per_step'], d['roofline'])"
timeout 300 python bench.py --steps 20 --warmup 5 --config infer --no-cpu-baseline > gpurun_out/r2_bench_fused_infer.json 2>> gpurun_out/r2_bench_fused.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fused_infer.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches_per_step'])"
grep -v -i warn gpurun_out/r2_bench_fused.err | tail -5
