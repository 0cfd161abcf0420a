"""Timing of eav_preproc_run for 42 subjects with EAV_PREPROC_GROUPS = 1, 2, 3, 6 (FIR of group i+1 overlapping the SOS of group i)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scipy.signal import butter
from eav_b200 import ops
from eav_b200.Dataload_eeg import decimation_taps, epoch_slots
S = 42
raw = torch.randn(S, 200, 30, 10000, device="cuda")
labels = []
for s in range(S):
    cls = np.random.default_rng(s).permutation(np.repeat(np.arange(10), 20))
    lab = np.zeros((10, 200)); lab[cls, np.arange(200)] = 1.0
    labels.append(lab)
slot = torch.from_numpy(np.stack([epoch_slots(l)[0] for l in labels])).cuda()
taps, sos = decimation_taps(5), butter(5, [0.5, 45], btype="bandpass", fs=100, output="sos")
eng = ops.PreprocEngine(S)
ep = torch.empty(S, 400, 30, 500, device="cuda")
out = {}
for exact in ("0", "1"):
    for g in ("1", "2", "3", "6", "14"):
        os.environ["EAV_PREPROC_GROUPS"] = g
        os.environ["EAV_SOS_EXACT"] = exact
        eng.run(raw, taps, sos, slot, 400, epochs=ep); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): eng.run(raw, taps, sos, slot, 400, epochs=ep)
        b.record(); torch.cuda.synchronize()
        out[f"exact{exact}_groups{g}"] = a.elapsed_time(b) / 3
print(json.dumps(out))
