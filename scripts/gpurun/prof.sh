#!/bin/bash
# ncu: full capture of the tensor-core kernels + launch list of one eager training step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"_tc_kernel" -c 5 -o gpurun_out/prof_r1_tc python scripts/profile_step.py --steps 1 --skip-preproc > gpurun_out/prof_tc.log 2>&1
tail -1 gpurun_out/prof_tc.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_step_launches_tc.csv python scripts/profile_step.py --steps 2 --skip-preproc > gpurun_out/prof_tc2.log 2>&1
tail -1 gpurun_out/prof_tc2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_bench_launches_tc.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -c 300 gpurun_out/b_ncu.log
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tc.py -m gpu -q -k "matches" 2>&1 | tail -4 > gpurun_out/r1_sanitizer_tc.txt
cat gpurun_out/r1_sanitizer_tc.txt
