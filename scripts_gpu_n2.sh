#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; tail -5 gpurun_out/bench_n2.err; head -c 600 gpurun_out/bench_n2.json; echo
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -3 gpurun_out/bench_ref.err; head -c 400 gpurun_out/bench_ref.json; echo
