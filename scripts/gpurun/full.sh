#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -c 3000 gpurun_out/bench_tc.json
