#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"fir_decimate" -c 2 -o gpurun_out/prof_fir python scripts/profile_step.py --steps 1 --subjects 8 --models 2 > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
