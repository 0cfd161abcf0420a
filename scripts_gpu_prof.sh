#!/bin/bash
# ncu: full capture of the two tensor-core kernels + launch list of one eager training step and of bench.py
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tconv_fwd_tc|tconv_bwd_dw_tc|tconv_wt_pack" -c 3 -o gpurun_out/prof_r1_tc python scripts/profile_step.py --steps 1 --skip-preproc > gpurun_out/prof_tc.log 2>&1
tail -2 gpurun_out/prof_tc.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_step_launches_tc.csv python scripts/profile_step.py --steps 2 --skip-preproc > gpurun_out/prof_tc2.log 2>&1
tail -1 gpurun_out/prof_tc2.log
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tc.py -m gpu -q -k "matches and not 500" 2>&1 | tail -6 > gpurun_out/r1_sanitizer_tc.txt
cat gpurun_out/r1_sanitizer_tc.txt
