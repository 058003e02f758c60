#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dataparallel.py -m gpu -q > gpurun_out/pytest_dp.log 2>&1
tail -15 gpurun_out/pytest_dp.log
