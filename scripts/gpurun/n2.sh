#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -12
for c in "nccl" "peer" "peer --graph"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/large_batch_sweep.py --batches 128,1024,8192 --collective $c 2> gpurun_out/sweep.err | grep '^\[' > gpurun_out/sweep_tmp.json; tail -2 gpurun_out/sweep.err | cut -c1-300; python -c "
import json; [print(r['collective'], r['cuda_graph'], r['global_batch'], r['bn'], round(r['ms_per_step'],4)) for r in json.load(open('gpurun_out/sweep_tmp.json'))]"
cp gpurun_out/sweep_tmp.json "gpurun_out/sweep_n2_$(echo $c | tr -d ' -').json"
done
