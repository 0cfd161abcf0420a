"""BASELINE.json configs[4]: large-batch synthetic EEGNet sweep, one model, batch split over the
ranks, NCCL all-reduce of BN sums + the flat gradient arena (eav_b200.data_parallel).
    python scripts/large_batch_sweep.py                      # 1 GPU
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/large_batch_sweep.py"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
from eav_b200.data_parallel import DataParallelEEGNet
from eav_b200.ops import EegnetDims

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="32,128,512,2048,8192")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--collective", default="auto")
ap.add_argument("--graph", action="store_true", help="replay the whole step as a CUDA graph (peer collective)")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
sd = EEGNet_tor(5).state_dict()
out = []
for GB in [int(b) for b in a.batches.split(",")]:
    if GB % world:
        continue
    B = GB // world
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn(B, 30, 500, generator=g, device=dev)
    y = torch.randint(0, 5, (B,), generator=g, device=dev)
    for mode in ("train", "eval"):
        dp = DataParallelEEGNet(EegnetDims(5), GB, lr=1e-5, state_dict=sd, bn_names=EEGNet_tor._BN_NAMES, collective=a.collective)
        for _ in range(3):
            dp.step(x, y, bn_train=mode == "train", graph=a.graph)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = dp.step(x, y, bn_train=mode == "train", graph=a.graph)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        out.append({"global_batch": GB, "world": world, "bn": mode, "ms_per_step": float(t), "samples_per_s": GB / float(t) * 1e3, "loss": float(loss), "collective": dp.collective, "cuda_graph": a.graph})
        del dp
if rank == 0:
    print(json.dumps(out))
if world > 1:
    torch.distributed.destroy_process_group()
