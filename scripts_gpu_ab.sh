#!/bin/bash
python scripts/kbench.py --stages dw_fwd,dw_bwd,pool1_fwd,pool1_bwd,tail_fwd,tail_bwd --reps 20
python -m pytest tests/test_gpu_eegnet.py -m gpu -q 2>&1 | tail -2
