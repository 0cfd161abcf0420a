#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_preproc.py tests/test_gpu_dropin.py -m gpu -q --timeout=900 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pre.csv python scripts/profile_step.py --steps 1 --subjects 42 --models 2 > gpurun_out/prof1.log 2>&1
grep -E "fir|sos" gpurun_out/launches_pre.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-stages > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'preprocess',d['preprocess']['value'],d['preprocess']['ms'],d['preprocess']['roofline']['frac'])"
