#!/bin/bash
# 2-GPU bench (torchrun) with and without the overlapped gradient exchange.  Needs a 2-GPU box: gpurun --gpus 2 -- bash scripts/gpu_call_ddp.sh
if [ "$(nvidia-smi -L | wc -l)" -lt 2 ]; then echo "needs 2 GPUs (gpurun --gpus 2)"; exit 1; fi
mkdir -p gpurun_out
for ov in 1 0 1 0; do
PGPD_DDP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('overlap=$ov', d['ms_per_step'], d['value'], d['cuda_graph'])"
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1gpu', d['ms_per_step'], d['value'])"
