#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_eegnet.py tests/test_gpu_tc.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
echo "== kbench"; timeout 100 python scripts/kbench.py --stages dw_bwd 2>&1 | tail -1
timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
