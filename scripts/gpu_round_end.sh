#!/bin/bash
# Round-end evidence run (under gpurun, 1 GPU): parity tests, smoke, bench at every BASELINE config, ncu launch list of the bench
# command, ncu --set full captures of the dominant kernels.  Outputs: gpurun_out/r2_final3_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2_final3_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_final3_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final3_bench.json 2> gpurun_out/r2_final3_bench.err; cut -c1-300 gpurun_out/r2_final3_bench.json
timeout 300 python bench.py --steps 20 --warmup 5 --classes 3 --no-cpu-baseline > gpurun_out/r2_final3_bench_k3.json 2>> gpurun_out/r2_final3_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --batch 128 --points 2048 --no-cpu-baseline > gpurun_out/r2_final3_bench_128x2048.json 2>> gpurun_out/r2_final3_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --config infer > gpurun_out/r2_final3_bench_infer.json 2>> gpurun_out/r2_final3_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --config tower > gpurun_out/r2_final3_bench_tower.json 2>> gpurun_out/r2_final3_bench.err
for f in k3 128x2048 infer tower; do python -c "
import json; d=json.load(open('gpurun_out/r2_final3_bench_$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches_per_step'])"; done
# launch list of the bench command (eager launches so that every kernel is visible; first step = warm-up is skipped by -s)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2_final3_launches.csv \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-graph > gpurun_out/r2_final3_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_l3_fwd_tc3|k_ka_tc|k_kb_tc|k_kf_tc|k_da2_sparse|k_dw3' -s 16 -c 8 -o gpurun_out/r2_final3_prof_train -f \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-graph > gpurun_out/r2_final3_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tower_fused_eval' -s 2 -c 1 -o gpurun_out/r2_final3_prof_fused -f \
    python bench.py --config tower --steps 2 --warmup 2 --no-cpu-baseline --no-graph > gpurun_out/r2_final3_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
