"""A/B check of the eval-mode fused block-1 backward (tconv_bwd_fused_tc_kernel) against the unfused pair
(dw_bwd + tconv_bwd_dw_tc): every gradient, several shapes, then the stage times at the bench size.
    python scripts/check_fused_bwd.py            # on a B200"""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eav_b200 import _lib, ops
from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
from eav_b200.trainer_core import SubjectBatchTrainer


def run(M, B, fused, n_rows=64, seed=0):
    os.environ["EAV_FUSE_BWD"] = "1" if fused else "0"
    os.environ["EAV_FUSE_FWD"] = "1" if fused else "0"
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M * n_rows, 30, 500, generator=g).cuda()
    y = torch.randint(0, 5, (M * n_rows,), generator=g).cuda()
    sds = []
    for m in range(M):
        torch.manual_seed(100 + m)
        mdl = EEGNet_tor(5)
        with torch.no_grad():                       # non-trivial BN state so the eval-mode affine matters
            for bn in (mdl.firstBN, mdl.depthwiseBN, mdl.separableBN):
                bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
                bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        sds.append(mdl.state_dict())
    core = SubjectBatchTrainer(mdl._dims, M, x, y, lr=1e-3, max_batch=B, use_graph=False)
    core.load_state_dicts(sds, EEGNet_tor._BN_NAMES)
    idx = (torch.arange(M).unsqueeze(1) * n_rows + torch.randperm(n_rows, generator=g)[:B].unsqueeze(0)).reshape(-1).int().cuda()
    p = core.program(B, False, "train")
    p.idx.copy_(idx)
    p.enqueue()
    torch.cuda.synchronize()
    return core.grads.clone(), p.loss.clone(), core


worst = 0.0
for M, B in ((1, 8), (3, 24), (2, 32), (5, 17)):
    ga, la, ca = run(M, B, True)
    gb, lb, cb = run(M, B, False)
    n, layout = ca.dims.param_layout()
    for name, off, shape in layout:
        k = int(np.prod(shape))
        a, b = ga[:, off:off + k].double(), gb[:, off:off + k].double()
        rel = float((a - b).norm() / b.norm().clamp_min(1e-30))
        worst = max(worst, rel)
        flag = "" if rel < 2e-5 else "   <-- MISMATCH"
        if flag or name in ("firstConv.weight", "firstBN.weight", "firstBN.bias", "depthwiseConv.weight"):
            print(f"M={M} B={B} {name:24s} rel-L2 {rel:.2e}{flag}")
    assert (la - lb).abs().max() < 1e-6 * lb.abs().max(), (la, lb)
print("worst rel-L2 fused vs unfused:", worst)
assert worst < 2e-5, worst

# stage times at the bench size
lib = _lib.load()
res = {}
for fused in (True, False):
    _, _, core = run(42, 32, fused, n_rows=64)
    p = core.program(32, False, "train")
    cfg = p.cfg()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for s_id in range(lib.eav_eegnet_stage_count()):
        name = lib.eav_eegnet_stage_name(s_id).decode()
        if name not in ("dw_fwd", "bn2_finalize", "pool1_fwd", "dw_bwd", "bn1_bwd_finalize", "bn2_bwd_finalize", "tconv_bwd_dw", "pool1_bwd"):
            continue
        def go():
            _lib.check(lib.eav_eegnet_run_stage(ctypes.byref(cfg), s_id, ops._ptr(core.x), ops._ptr(p.idx), ops._ptr(core.params),
                                                ops._ptr(core.bn_state), None, None, ops._ptr(p.out), ops._ptr(p.dout),
                                                ops._ptr(core.grads), ops._ptr(core.workspace), core.ws_bytes, st), "stage")
        go(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            go()
        b.record(); torch.cuda.synchronize()
        out[name] = a.elapsed_time(b) / 10
    res["fused" if fused else "unfused"] = out
    del core
print(json.dumps(res))
