#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -2
timeout 300 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "matches and (shape4 or shape1)" 2>&1 | grep -E "Barrier error|Device Frame: eav::<unnamed>|ERROR SUMMARY|passed|failed" | sed -E "s/\+0x[0-9a-f]+//" | sort | uniq -c | sort -rn | head -8
echo "== kbench"; timeout 100 python scripts/kbench.py --stages tconv_bwd_dw 2>&1 | tail -1
