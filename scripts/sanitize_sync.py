"""synccheck discriminator for the 'Barrier error: Missing init' report (profiles/r1_sanitizer_tc.txt):
(1) eav_tc_probe -- 128 threads, ONE mbarrier, initialised by thread 0 before a __syncthreads and completed only by a
tcgen05.commit arrival -- and (2) the fused eval-mode backward with enough rows per unit that its producers wait on
`bar_empty` (also completed only by tcgen05.commit).  If the trivially correct probe kernel draws the same report, the
report is about tcgen05.commit arrivals not being modelled by the tool, not about the kernels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eav_b200 import ops
from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
from eav_b200.trainer_core import SubjectBatchTrainer
which = sys.argv[1]
if which == "probe":
    img = torch.zeros(4096, device="cuda")
    d, cyc = ops.tc_probe(img, 128, 32, 2, 1, (0, 16, 128, 0, 32), (8192, 128, 256, 0, 32))
    print("probe ok", float(d.abs().sum()), cyc)
else:
    M, B = 1, 32
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 30, 500, generator=g).cuda(); y = torch.randint(0, 5, (64,), generator=g).cuda()
    torch.manual_seed(1); mdl = EEGNet_tor(5)
    core = SubjectBatchTrainer(mdl._dims, M, x, y, lr=1e-3, max_batch=B, use_graph=False)
    core.load_state_dicts([mdl.state_dict()], EEGNet_tor._BN_NAMES)
    idx = torch.arange(B).int().cuda()
    print(core.train_step(idx, bn_train=False).cpu().tolist())
