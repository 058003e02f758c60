#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "exit=$?" >> gpurun_out/bench_2gpu.log
cut -c1-300 gpurun_out/bench_2gpu.log; grep -v "OMP_NUM\|\*\*\*" gpurun_out/bench_2gpu.err | tail -5
