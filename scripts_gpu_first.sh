#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'eval',d.get('eval_bn_step'),'e2e',d['e2e']['value'])
print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','peak','frac','share_of_step')})
print({k:round(v,3) for k,v in d['stage_ms'].items() if v>0.05})
print('preprocess',d['preprocess']['value'],d['preprocess']['ms'],d['preprocess']['roofline']['frac'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pre.csv python scripts/profile_step.py --steps 1 --subjects 42 --models 2 > gpurun_out/prof1.log 2>&1
grep -E "fir|sos" gpurun_out/launches_pre.csv | awk -F'","' '{print $5, $NF}' | tail -8
