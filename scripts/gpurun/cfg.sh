#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/train_all_subjects.py --subjects 42 --epochs 4 --lr 1e-3 --separable > gpurun_out/e2e_42subjects_n$N.json 2> gpurun_out/e2e.err; tail -2 gpurun_out/e2e.err; cat gpurun_out/e2e_42subjects_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/large_batch_sweep.py > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; cat gpurun_out/sweep_n$N.json
