#!/bin/bash
mkdir -p gpurun_out
# (1) launch list of the bench command itself
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/prof1.log
tail -2 gpurun_out/prof1.log
# (2) one eager train step + one preprocessing pass, every kernel its own launch
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python scripts/profile_step.py --steps 2 --subjects 42 > gpurun_out/prof1b.log 2>&1
# (3) full-set capture of the top kernels (1 launch each, second pass of each)
ncu --set full --clock-control none --import-source on -k regex:"tconv_bwd_dw|dw_bwd_kernel|fir_decimate|sos_kernel|sepconv|tconv_fwd|dw_fwd" -s 3 -c 10 -o gpurun_out/prof_r1 python scripts/profile_step.py --steps 1 --subjects 42 > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
