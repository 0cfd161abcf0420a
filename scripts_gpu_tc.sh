#!/bin/bash
# tensor-core temporal conv: parity tests, then stage timings for both paths
timeout 300 python -m pytest tests/test_gpu_eegnet.py -m gpu -x -q 2>&1 | tail -15
echo "== kbench tc"; timeout 100 python scripts/kbench.py --stages tconv_fwd,tconv_bwd_dw 2>&1 | tail -4
