#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
timeout 400 python scripts/ingest_bench.py --subjects 4 2>&1 | tail -2 | tee gpurun_out/ingest_bench.json
