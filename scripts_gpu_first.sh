#!/bin/bash
# first GPU contact: run all gpu tests without -x and keep the log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
