#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -2
echo "== kbench tc"; timeout 100 python scripts/kbench.py --stages sepconv_fwd,sepconv_bwd_dx 2>&1 | tail -2
