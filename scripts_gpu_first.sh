#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
