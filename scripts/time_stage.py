"""Times single stages of the eval-mode EEGNet step at the bench size (42 models x 32) through eav_eegnet_run_stage.
    python scripts/time_stage.py dw_fwd tconv_fwd        # on a B200; environment switches (EAV_*) apply"""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eav_b200 import _lib, ops
from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
from eav_b200.trainer_core import SubjectBatchTrainer

names = sys.argv[1:] or ["dw_fwd"]
M, B, n_rows = int(os.environ.get("M", 42)), 32, 64
g = torch.Generator().manual_seed(0)
x = torch.randn(M * n_rows, 30, 500, generator=g).cuda()
y = torch.randint(0, 5, (M * n_rows,), generator=g).cuda()
torch.manual_seed(1)
mdl = EEGNet_tor(5)
core = SubjectBatchTrainer(mdl._dims, M, x, y, lr=1e-3, max_batch=B, use_graph=False)
core.load_state_dicts([mdl.state_dict()] * M, EEGNet_tor._BN_NAMES)
p = core.program(B, False, "train")
p.idx.copy_((torch.arange(M).unsqueeze(1) * n_rows + torch.arange(B).unsqueeze(0)).reshape(-1).int().cuda())
p.enqueue(); torch.cuda.synchronize()
lib = _lib.load()
cfg = p.cfg()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for s_id in range(lib.eav_eegnet_stage_count()):
    name = lib.eav_eegnet_stage_name(s_id).decode()
    if name not in names:
        continue
    def go():
        _lib.check(lib.eav_eegnet_run_stage(ctypes.byref(cfg), s_id, ops._ptr(core.x), ops._ptr(p.idx), ops._ptr(core.params),
                                            ops._ptr(core.bn_state), None, None, ops._ptr(p.out), ops._ptr(p.dout),
                                            ops._ptr(core.grads), ops._ptr(core.workspace), core.ws_bytes, st), "stage")
    go(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); go(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    out[name] = round(tot / 10, 4)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("EAV_")}, "M": M, "ms": out}))
