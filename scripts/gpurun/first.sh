#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'eval',d.get('eval_bn_step'),'e2e',d['e2e']['value'])
r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','peak','frac','peak_register_operands','frac_of_register_operand_peak','share_of_step')}, r['whole_step'])
print({k:round(v,3) for k,v in d['stage_ms'].items() if v>0.04})
PY
