"""Time individual EEGNet kernel stages (42 models x B=32) with CUDA events -- quick A/B driver."""
import argparse, ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eav_b200 import _lib, ops
from eav_b200.trainer_core import SubjectBatchTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--models", type=int, default=42)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--stages", default="tconv_fwd,tconv_bwd_dw,sepconv_fwd,sepconv_bwd_dx,sepconv_bwd_dw,dw_fwd,dw_bwd")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--bn", default="train")
a = ap.parse_args()
dev = torch.device("cuda", 0)
M, B = a.models, a.batch
x = torch.randn(M * 280, 30, 500, device=dev)
y = torch.randint(0, 5, (M * 280,), device=dev)
tr = SubjectBatchTrainer(ops.EegnetDims(5), M, x, y, lr=1e-5, max_batch=B, use_graph=False)
tr.params.normal_(0, 0.05)
idx = (torch.stack([torch.randperm(280)[:B] for _ in range(M)]) + torch.arange(M).unsqueeze(1) * 280).reshape(-1).int().to(dev)
tr.train_step(idx, bn_train=a.bn == "train")
torch.cuda.synchronize()
lib = _lib.load()
prog = tr.program(B, a.bn == "train", "train")
cfg = prog.cfg()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
names = {lib.eav_eegnet_stage_name(i).decode(): i for i in range(lib.eav_eegnet_stage_count())}
for nm in a.stages.split(","):
    sid = names[nm]
    def go():
        _lib.check(lib.eav_eegnet_run_stage(ctypes.byref(cfg), sid, ops._ptr(tr.x), ops._ptr(prog.idx), ops._ptr(tr.params),
                                            ops._ptr(tr.bn_state), None, None, ops._ptr(prog.out), ops._ptr(prog.dout),
                                            ops._ptr(tr.grads), ops._ptr(tr.workspace), tr.ws_bytes, st), nm)
    go(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps): go()
    e1.record(); torch.cuda.synchronize()
    print(f"{nm:18s} {e0.elapsed_time(e1) / a.reps:8.4f} ms")
