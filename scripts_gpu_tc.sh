#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_eegnet.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -3
echo "== kbench tc"; timeout 100 python scripts/kbench.py --stages sepconv_fwd,sepconv_bwd_dx,sepconv_bwd_dw 2>&1 | tail -3
