#!/bin/bash
for b in 3 2; do echo "minb $b"; EAV_TW_MINB=$b python scripts/kbench.py --stages tconv_bwd_dw --reps 20; done
python -m pytest tests/test_gpu_eegnet.py -m gpu -q 2>&1 | tail -1
EAV_TW_MINB=2 python -m pytest tests/test_gpu_eegnet.py -m gpu -q 2>&1 | tail -1
