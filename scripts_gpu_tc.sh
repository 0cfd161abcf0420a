#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_eegnet.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -2
echo "== kbench"; timeout 100 python scripts/kbench.py --stages tail_bwd,bn3_bwd_apply,pool1_fwd,pool1_bwd,tail_fwd,dw_fwd 2>&1 | tail -6
