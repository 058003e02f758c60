#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "eval or inference or shipped or simt_flag or large_magnitude or nan or reference_error" > gpurun_out/r2_pytest_fused.log 2>&1; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_fused.log | tail -15
for c in tower infer; do timeout 300 python bench.py --steps 20 --warmup 5 --config $c --no-cpu-baseline > gpurun_out/r2_bench_fused_$c.json 2> gpurun_out/r2_bench_fused.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_fused_$c.json')); print('$c', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done
grep -v -i warn gpurun_out/r2_bench_fused.err | tail -5
