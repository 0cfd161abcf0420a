#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; tail -3 gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/large_batch_sweep.py --batches 128,1024,8192 > gpurun_out/sweep_n2.json 2> gpurun_out/sweep.err; tail -1 gpurun_out/sweep.err; cat gpurun_out/sweep_n2.json
