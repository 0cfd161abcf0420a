#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"dw_bwd_kernel|dw_fwd_kernel|sepconv_bwd_dw|pool1|tail_" -c 7 -o gpurun_out/prof_r1b python scripts/profile_step.py --steps 1 --skip-preproc > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
