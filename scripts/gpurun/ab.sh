#!/bin/bash
python -m pytest tests/test_gpu_preproc.py tests/test_gpu_dropin.py::test_dataload_eeg_dropin_vs_oracle -m gpu -q 2>&1 | tail -1
for g in 1 2 3 6; do echo "groups $g"; EAV_PREPROC_GROUPS=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-stages 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['preprocess']['ms'], d['preprocess']['value'], d['preprocess']['roofline']['frac'])"; done
