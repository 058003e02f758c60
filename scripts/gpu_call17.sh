#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q > gpurun_out/pytest_graph.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "exit=$?" >> gpurun_out/bench_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/bench_2gpu_eager.log 2>> gpurun_out/bench_2gpu.err
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/pytest_graph.log; tail -2 gpurun_out/smoke.log; cut -c1-420 gpurun_out/bench_2gpu.log; cut -c1-200 gpurun_out/bench_2gpu_eager.log; tail -5 gpurun_out/bench_2gpu.err; cut -c1-250 gpurun_out/bench_ref.log
