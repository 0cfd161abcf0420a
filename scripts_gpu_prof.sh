#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"sepconv_tc_kernel" -c 2 -o gpurun_out/prof_r1_sct python scripts/profile_step.py --steps 1 --skip-preproc > gpurun_out/prof_sct.log 2>&1
tail -2 gpurun_out/prof_sct.log
