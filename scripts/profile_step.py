"""Small driver for ncu: one preprocessing pass (S subjects) and a few EAGER (non-graph)
training steps of M models x B=32, so every kernel shows up as its own launch."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eav_b200 import ops
from eav_b200.trainer_core import SubjectBatchTrainer
from eav_b200.Dataload_eeg import decimation_taps, epoch_slots
from scipy.signal import butter

ap = argparse.ArgumentParser()
ap.add_argument("--models", type=int, default=42)
ap.add_argument("--subjects", type=int, default=8)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--bn", default="train")
ap.add_argument("--skip-preproc", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
if not a.skip_preproc:
    S = a.subjects
    raw = torch.randn(S, 200, 30, 10000, device=dev)
    lab = np.zeros((10, 200)); lab[np.repeat(np.arange(10), 20), np.arange(200)] = 1
    slot = torch.from_numpy(np.stack([epoch_slots(lab)[0]] * S)).to(dev)
    eng = ops.PreprocEngine(S, device=dev)
    for _ in range(2):
        eng.run(raw, decimation_taps(5), butter(5, [0.5, 45], btype="bandpass", fs=100, output="sos"), slot, 400)
    torch.cuda.synchronize()
    del raw, eng
M, B = a.models, 32
x = torch.randn(M * 280, 30, 500, device=dev)
y = torch.randint(0, 5, (M * 280,), device=dev)
tr = SubjectBatchTrainer(ops.EegnetDims(5), M, x, y, lr=1e-5, max_batch=B, use_graph=False)
tr.params.normal_(0, 0.05)
_, layout = tr.dims.param_layout()
for name, off, shape in layout:
    if name.endswith("BN.weight"):
        tr.params[:, off:off + int(np.prod(shape))] = 1.0
idx = (torch.stack([torch.randperm(280)[:B] for _ in range(M)]) + torch.arange(M).unsqueeze(1) * 280).reshape(-1).int().to(dev)
for _ in range(a.steps):
    tr.train_step(idx, bn_train=a.bn == "train")
torch.cuda.synchronize()
print("done")
