#!/bin/bash
echo "== kbench all stages in order (tc)"; timeout 100 python scripts/kbench.py --stages pool1_bwd,bn2_bwd_reduce,bn2_bwd_finalize,dw_bwd,bn1_bwd_finalize,tconv_bwd_dw,dw_bwd 2>&1 | tail -8
echo "== bench ffma"; EAV_TCONV=ffma timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v,3) for k,v in d['stage_ms'].items() if v>0.05})"
echo "== bench tc"; timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v,3) for k,v in d['stage_ms'].items() if v>0.05})"
