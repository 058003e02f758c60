#!/bin/bash
# ncu --set full of the backward kernels (one launch each) of a train step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_da2_sparse|k_dw3|k_ka_tc|k_kb_tc' -s 8 -c 4 -o gpurun_out/r2_final2_prof_bwd -f \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-graph > gpurun_out/r2_final2_ncu_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -3
