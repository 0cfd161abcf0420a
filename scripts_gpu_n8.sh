#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['preprocess']['value'], d['clocks'])"
if [ "$2" == "full" ]; then
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/train_all_subjects.py --subjects 42 --epochs 4 --lr 1e-3 --separable > gpurun_out/e2e_42subjects_n$N.json 2> gpurun_out/e2e.err; tail -1 gpurun_out/e2e.err; cat gpurun_out/e2e_42subjects_n$N.json
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/large_batch_sweep.py --batches 128,1024,8192 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep.err; tail -1 gpurun_out/sweep.err; cat gpurun_out/sweep_n$N.json
fi
