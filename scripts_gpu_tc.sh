#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -2
echo "== kbench"; timeout 100 python scripts/kbench.py --stages tconv_bwd_dw,tconv_fwd 2>&1 | tail -2
