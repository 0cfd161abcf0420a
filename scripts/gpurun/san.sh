#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "=== $tool"
  timeout 500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "matches and (shape3 or shape4 or shape1)" 2>&1 | tail -5
done > gpurun_out/r1_sanitizer_tc.txt 2>&1
cat gpurun_out/r1_sanitizer_tc.txt
