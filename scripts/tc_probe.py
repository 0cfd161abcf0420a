"""Pins the tcgen05 no-swizzle operand address maps (eav_b200/csrc/tc_common.cuh) on the device and times
tcgen05.mma.kind::tf32 for the operand shapes the temporal-conv kernels use.

    python scripts/tc_probe.py            # prints one line per case, writes gpurun_out/tc_probe.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eav_b200 import ops  # noqa: E402


def addr_k_major(r, k, off, lbo, sbo):
    return off + (r // 8) * sbo + (r % 8) * 16 + (k // 4) * lbo + (k % 4) * 4


def addr_mn_major(r, k, off, lbo, sbo):
    return off + (r // 4) * sbo + (r % 4) * 4 + (k % 8) * 16 + (k // 8) * lbo


def operand(image, rows, desc, kstep, swap=False):
    off, lbo, sbo, major, step = desc
    if swap:
        lbo, sbo = sbo, lbo
    r = np.arange(rows)[:, None]
    k = np.arange(8)[None, :]
    fn = addr_mn_major if major else addr_k_major
    a = fn(r, k, off + kstep * step, lbo, sbo)
    return image[a // 4]


def model(image, M, N, ksteps, a, b, swap_a=False, swap_b=False):
    d = np.zeros((M, N), np.float64)
    for ks in range(ksteps):
        d += operand(image, M, a, ks, swap_a).astype(np.float64) @ operand(image, N, b, ks, swap_b).astype(np.float64).T
    return d


def main():
    rng = np.random.default_rng(0)
    nfl = 48 * 1024
    image = (rng.integers(-8, 9, nfl) / 4.0).astype(np.float32)
    dev = torch.from_numpy(image).cuda()
    B0 = 64 * 1024
    out = {"cases": [], "timing": []}

    def packed_b(N):   # K-major operand stored as [n/8][k/4][8][4]
        return (B0, 128, 256, 0, N * 32)

    cases = [
        ("fwd: A raw K-major {LBO16,SBO128}, B packed K-major", 128, 32, 4, (0, 16, 128, 0, 32), packed_b(32)),
        ("fwd, A start +592 B", 128, 32, 4, (592, 16, 128, 0, 32), packed_b(32)),
        ("fwd, N=64", 128, 64, 4, (0, 16, 128, 0, 32), packed_b(64)),
        ("bwd: A raw MN-major {SBO16}, B rows MN-major {SBO=pitch 2064}", 128, 32, 4, (0, 128, 16, 1, 128),
         (B0, 128, 2064, 1, 128)),
        ("bwd, A start +1216 B (third M tile)", 128, 32, 4, (1216, 128, 16, 1, 128), (B0, 128, 2064, 1, 128)),
        ("bwd, M=64", 64, 32, 4, (0, 128, 16, 1, 128), (B0, 128, 2064, 1, 128)),
    ]
    for name, M, N, ks, a, b in cases:
        rec = {"name": name, "M": M, "N": N}
        try:
            d, cyc = ops.tc_probe(dev, M, N, ks, 1, a, b)
            d = d.cpu().numpy().astype(np.float64)
            rec["cycles"] = cyc
            for sa in (False, True):
                for sb in (False, True):
                    ref = model(image, M, N, ks, a, b, sa, sb)
                    rows = d[:M] if M == 128 else d[:64]
                    rec[f"maxerr_swapA{int(sa)}_swapB{int(sb)}"] = float(np.abs(rows - ref).max())
            if M == 64:   # where do the 64 rows land in the 128 lanes?
                ref = model(image, M, N, ks, a, b)
                rec["m64_lanes_0_63"] = float(np.abs(d[:64] - ref).max())
                rec["m64_lanes_0_31_64_95"] = float(np.abs(np.concatenate([d[0:32], d[64:96]]) - ref).max())
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)
        print(json.dumps(rec), flush=True)
        out["cases"].append(rec)

    for N in (32, 64, 96, 128, 256):
        for label, a in (("A K-major raw", (0, 16, 128, 0, 32)), ("A MN-major raw", (0, 128, 16, 1, 128)),
                         ("A K-major packed", (0, 128, 256, 0, 0))):
            rec = {"timing": label, "N": N}
            try:
                if N % 32:
                    continue
                reps, ks = 64, 16
                b = (B0, 128, 256, 0, 0)
                _, c1 = ops.tc_probe(dev, 128, N, ks, reps, a, b)
                _, c2 = ops.tc_probe(dev, 128, N, ks, 2 * reps, a, b)
                rec["cycles_per_mma"] = (c2 - c1) / (reps * ks)
            except Exception as e:  # noqa: BLE001
                rec["error"] = repr(e)
            print(json.dumps(rec), flush=True)
            out["timing"].append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tc_probe.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
