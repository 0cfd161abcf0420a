#!/bin/bash
echo "split_tail=0"; EAV_SEP_SPLIT_TAIL=0 python scripts/kbench.py --stages sepconv_fwd,sepconv_bwd_dx --reps 20
echo "split_tail=1"; EAV_SEP_SPLIT_TAIL=1 python scripts/kbench.py --stages sepconv_fwd,sepconv_bwd_dx --reps 20
for s in 2 4 6 8 16; do echo "dw splits $s"; EAV_SEPDW_SPLITS=$s python scripts/kbench.py --stages sepconv_bwd_dw --reps 20; done
