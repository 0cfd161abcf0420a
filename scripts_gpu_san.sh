#!/bin/bash
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_preproc.py tests/test_gpu_eegnet.py tests/test_gpu_dropin.py::test_trainer_uni_loop_matches_reference -m gpu -q -x --timeout=2300 > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_preproc.py::test_small_case_vs_reference_golden "tests/test_gpu_eegnet.py::test_tor_fwd_bwd_vs_reference[train-b8]" "tests/test_gpu_eegnet.py::test_many_models_one_launch_equals_models_one_by_one[True]" -m gpu -q -x --timeout=1400 > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_racecheck.log | tail -5
