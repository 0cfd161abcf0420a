"""Decodes, on the device, which shared-memory byte a tcgen05.mma.kind::tf32 reads for operand element (r, k)
under the MN-major no-swizzle descriptor, by multiplying an index-valued operand with a one-hot one.
Also times the MMA when it rotates over several independent TMEM accumulators.

    python scripts/tc_decode.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eav_b200 import ops  # noqa: E402

B0 = 64 * 1024


def main():
    nfl = 24 * 1024
    image = np.zeros(nfl, np.float32)
    image[:2048] = np.arange(2048)
    # one-hot K-major packed operand at B0: elem(r, k) = 1 if r == k
    for r in range(8):
        k = r
        image[(B0 + (k // 4) * 128 + (r % 8) * 16 + (k % 4) * 4) // 4] = 1.0
    dev = torch.from_numpy(image).cuda()
    onehot = (B0, 128, 256, 0, 0)
    out = {}
    SW128_32B = 1 << 29     # layout_type = 1 in descriptor bits 61..63

    def swz(a):
        return a ^ (((a >> 7) & 3) << 5)

    def decode(label, desc, bits, as_a=True):
        if as_a:    # operand under test is A, decoded through one-hot B:  D[m][n<8] = A[m][k=n]
            d, _ = ops.tc_probe(dev, 128, 32, 1, 1, desc, onehot, a_bits=bits)
            tab = (d.cpu().numpy()[:, :8] * 4).astype(int)
        else:       # operand under test is B, decoded through one-hot A:  D[m<8][n] = B[n][k=m]
            d, _ = ops.tc_probe(dev, 128, 32, 1, 1, onehot, desc, b_bits=bits)
            tab = (d.cpu().numpy()[:8, :].T * 4).astype(int)
        print(f"--- {label} {'A' if as_a else 'B'} desc={desc} bits={bits:#x}: byte offset of (r, k); rows r, cols k")
        for r in list(range(0, 12)) + [31, 32, 33, 64, 96, 127]:
            if r < tab.shape[0]:
                print(f"  r={r:3d}: {tab[r].tolist()}")
        out[f"{label}_{'A' if as_a else 'B'}_{desc}_{bits}"] = tab.tolist()
        return tab

    t = decode("sanity K-major none", (0, 16, 128, 0, 0), 0)
    off, lbo, sbo = 0, 128, 512
    for (off, lbo, sbo) in ((0, 128, 512), (128, 128, 512), (256, 128, 512), (0, 2048, 512), (0, 2048, 1024)):
        for as_a in (True, False):
            tab = decode("MN SW128_32B", (off, lbo, sbo, 1, 0), SW128_32B, as_a)
            R = tab.shape[0]
            r = np.arange(R)[:, None]
            k = np.arange(8)[None, :]
            hyp = swz(off + (r // 32) * lbo + (k // 4) * sbo + (k % 4) * 128 + (r % 32) * 4)
            hyp2 = off + ((r // 32) * lbo + (k // 4) * sbo + (k % 4) * 128 + (r % 32) * 4 ^ ((k % 4) << 5))
            print("   matches absolute-address swizzle:", bool((tab == hyp).all()),
                  " matches start-relative swizzle:", bool((tab == hyp2).all()))
    tim = []
    for N in (32, 64, 128, 256):
        for n_acc in (1, 2, 3, 4, 8):
            if N * n_acc > 512:
                continue
            _, c1 = ops.tc_probe(dev, 128, N, 16, 64, (0, 16, 128, 0, 32), onehot, n_acc=n_acc)
            _, c2 = ops.tc_probe(dev, 128, N, 16, 128, (0, 16, 128, 0, 32), onehot, n_acc=n_acc)
            rec = {"N": N, "n_acc": n_acc, "cycles_per_mma": (c2 - c1) / (64 * 16)}
            print(json.dumps(rec), flush=True)
            tim.append(rec)
    out["timing"] = tim
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tc_decode.json", "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
