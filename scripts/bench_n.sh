#!/bin/bash
# bench.py on N GPUs of this box, launched as the driver launches it; writes gpurun_out/r2_bench_n$N.json
N=${1:-2}; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 90 --warmup 5 "$@" 2>gpurun_out/n$N.err | tail -1 > gpurun_out/r2_bench_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_parity',{}).get('ok'))"
