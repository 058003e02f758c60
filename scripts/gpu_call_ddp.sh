#!/bin/bash
mkdir -p gpurun_out
for ov in 1 0 1 0; do
PGPD_DDP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('overlap=$ov', d['ms_per_step'], d['value'], d['cuda_graph'])"
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1gpu', d['ms_per_step'], d['value'])"
