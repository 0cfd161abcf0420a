#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
for c in "nccl" "peer --graph"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/large_batch_sweep.py --batches 128,1024,8192 --collective $c 2> gpurun_out/sweep.err | grep '^\[' > gpurun_out/sweep_tmp.json; tail -2 gpurun_out/sweep.err | cut -c1-300; python -c "
import json; [print(r['collective'], r['cuda_graph'], r['global_batch'], r['bn'], round(r['ms_per_step'],4)) for r in json.load(open('gpurun_out/sweep_tmp.json'))]"
cp gpurun_out/sweep_tmp.json "gpurun_out/sweep_n${N}_$(echo $c | tr -d ' -').json"
done
