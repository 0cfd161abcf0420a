"""Tiny workload for compute-sanitizer: one train-mode and one eval-mode step of M=2 models x B=8 (every kernel of
the EEGNet path incl. the tcgen05 kernels and the fused eval-mode backward), a ShallowConvNet step and a small
preprocessing pass (both SOS paths).
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_step.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scipy.signal import butter
from eav_b200 import ops
from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
from eav_b200.Dataload_eeg import decimation_taps
from eav_b200.trainer_core import SubjectBatchTrainer
from eav_b200.Transformer_torch.Transformer_EEG import ShallowConvNet

M, B = 2, 8
g = torch.Generator().manual_seed(0)
x = torch.randn(M * 16, 30, 500, generator=g).cuda()
y = torch.randint(0, 5, (M * 16,), generator=g).cuda()
torch.manual_seed(1)
mdl = EEGNet_tor(5)
core = SubjectBatchTrainer(mdl._dims, M, x, y, lr=1e-3, max_batch=B, use_graph=False)
core.load_state_dicts([mdl.state_dict()] * M, EEGNet_tor._BN_NAMES)
idx = (torch.arange(M).unsqueeze(1) * 16 + torch.arange(B).unsqueeze(0)).reshape(-1).int().cuda()
for bn_train in (True, False):
    loss = core.train_step(idx, bn_train=bn_train)
    torch.cuda.synchronize()
    print("eegnet step bn_train=%s loss %s" % (bn_train, loss.cpu().tolist()))
if "--no-shallow" not in sys.argv:
    net = ShallowConvNet(5, num_layers=2).cuda().train()
    xs = torch.randn(2, 1, 30, 500, generator=g).cuda()
    out = net(xs)
    torch.nn.functional.cross_entropy(out, torch.tensor([0, 1]).cuda()).backward()
    torch.cuda.synchronize()
    print("shallow step ok")
raw = torch.randn(1, 6, 30, 10000, generator=g).cuda()
slot = torch.tensor([[0, -1, 1, 2, -1, 3]], dtype=torch.int32).cuda()
eng = ops.PreprocEngine(1, n_trials=6)
for exact in ("0", "1"):
    os.environ["EAV_SOS_EXACT"] = exact
    ep = eng.run(raw, decimation_taps(5), butter(5, [0.5, 45], btype="bandpass", fs=100, output="sos"), slot, 16)
    torch.cuda.synchronize()
    print("preproc exact=%s ok" % exact, float(ep.abs().mean()))
