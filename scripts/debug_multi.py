import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gpu_util as U
from eav_b200.ops import EegnetDims, EegnetEngine
g = np.load(os.path.join(ROOT, "tests/golden/eegnet_tor_b8.npz"))
sd0 = U.init_from_golden(g)
dims = EegnetDims(5)
gen = torch.Generator().manual_seed(77)
M, B = 12, 32
sds = []
for m in range(M):
    sd = {k: (v.clone() if v.dtype != torch.float32 else v + 0.03 * torch.randn(v.shape, generator=gen)) for k, v in sd0.items()}
    for bnn in U.TOR_BN: sd[bnn + ".running_var"] = sd[bnn + ".running_var"].abs() + 0.5
    sds.append(sd)
x = torch.randn(M * B, 30, 500, generator=gen).cuda()
for train in (False, True):
    params, bn = U.pack_params(dims, sds), U.pack_bn(dims, sds)
    m1 = (torch.rand(M * B, 64, 125, generator=gen) > 0.5).to(torch.uint8).cuda()
    m2 = (torch.rand(M * B, 64, 15, generator=gen) > 0.5).to(torch.uint8).cuda()
    eng = EegnetEngine(dims, M, B)
    out = eng.forward(x, params, bn, bn_train=train, mask1=m1 if train else None, mask2=m2 if train else None).clone()
    saved = {k: eng.saved(k).clone() for k in ("y1", "y2", "d1", "y3", "feat")}
    one = EegnetEngine(dims, 1, B)
    for m in (0, 1, 5, 11):
        sl = slice(m * B, (m + 1) * B)
        p1, b1 = U.pack_params(dims, [sds[m]]), U.pack_bn(dims, [sds[m]])
        o1 = one.forward(x[sl].contiguous(), p1, b1, bn_train=train, mask1=m1[sl].contiguous() if train else None, mask2=m2[sl].contiguous() if train else None)
        errs = {k: U.rel_max(saved[k][sl].cpu().numpy(), one.saved(k).cpu().numpy()) for k in saved}
        print("train" if train else "eval", "model", m, {k: f"{v:.1e}" for k, v in errs.items()}, f"out {U.rel_max(out[sl].cpu().numpy(), o1.cpu().numpy()):.1e}")
