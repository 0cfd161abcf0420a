#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python scripts/kbench.py --reps 20
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'eval',d.get('eval_bn_step'),'e2e',d['e2e']['value'])
print({k:round(v,3) for k,v in d['stage_ms'].items() if v>0.04})
print('preprocess',d['preprocess']['value'],d['preprocess']['ms'],d['preprocess']['roofline']['frac'])
PY
