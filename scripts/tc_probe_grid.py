import os, sys, json, numpy as np, torch
sys.path.insert(0, "/root/repo")
from eav_b200 import ops
image = np.zeros(20*1024, np.float32); dev = torch.from_numpy(image).cuda()
B0 = 64*1024
onehot = (B0, 128, 256, 0, 0)
for grid in (1, 148, 296):
    os.environ["EAV_TC_PROBE_GRID"] = str(grid)
    for M in (128, 64):
        for N in (32, 64, 128):
            _, c1 = ops.tc_probe(dev, M, N, 16, 64, (0, 16, 128, 0, 32), onehot, n_acc=2)
            _, c2 = ops.tc_probe(dev, M, N, 16, 128, (0, 16, 128, 0, 32), onehot, n_acc=2)
            print(json.dumps({"grid": grid, "M": M, "N": N, "cycles_per_mma": (c2-c1)/1024}), flush=True)
