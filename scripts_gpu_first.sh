#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
