#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python scripts/profile_step.py --steps 2 > gpurun_out/prof1.log 2>&1
tail -3 gpurun_out/prof1.log
ncu --set full --clock-control none --import-source on -k regex:"tconv_bwd_dw|tconv_fwd|dw_bwd_kernel|fir_decimate|sos_kernel" -c 8 -o gpurun_out/prof_r1 python scripts/profile_step.py --steps 1 --subjects 4 > gpurun_out/prof2.log 2>&1
tail -3 gpurun_out/prof2.log
ls -la gpurun_out
