#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_eegnet.py tests/test_gpu_tc.py tests/test_gpu_dropin.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
echo "== kbench"; timeout 100 python scripts/kbench.py --stages pool1_fwd,pool1_bwd 2>&1 | tail -2
