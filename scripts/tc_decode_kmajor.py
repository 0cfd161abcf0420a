"""Decodes the K-major SWIZZLED tf32 operand address maps (SWIZZLE_32B / 64B / 128B) on the device, for start
addresses shifted by whole rows: the information a J = 8 / 16 / 32 variant of the Toeplitz trick needs.

    python scripts/tc_decode_kmajor.py      # prints one line per case, writes gpurun_out/tc_decode_kmajor.json
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
    image = np.zeros(24 * 1024, np.float32)
    image[:2048] = np.arange(2048)
    for r in range(8):
        image[(B0 + (r // 4) * 128 + (r % 8) * 16 + (r % 4) * 4) // 4] = 1.0      # one-hot K-major packed operand
    dev = torch.from_numpy(image).cuda()
    onehot = (B0, 128, 256, 0, 0)
    out = []
    # (name, layout_type, row pitch bytes, swizzle: XOR bits [4, 4+nb) with bits [7, 7+nb))
    for name, lt, pitch, nb in (("SWIZZLE_32B", 6, 32, 1), ("SWIZZLE_64B", 4, 64, 2), ("SWIZZLE_128B", 2, 128, 3)):
        for off in (0, pitch, 2 * pitch, 8 * pitch, 1024 + 3 * pitch):
            sbo = 8 * pitch
            d, _ = ops.tc_probe(dev, 128, 32, 1, 1, (off, 16, sbo, 0, 0), onehot, a_bits=lt << 29)
            tab = (d.cpu().numpy()[:, :8] * 4).astype(int)
            r, k = np.arange(128)[:, None], np.arange(8)[None, :]
            a = off + (r // 8) * sbo + (r % 8) * pitch + k * 4
            mask = (1 << nb) - 1
            absolute = a ^ (((a >> 7) & mask) << 4)
            rel = off + ((a - off) ^ ((((a - off) >> 7) & mask) << 4))
            rows_ok = a.max(axis=1) + 16 < 2048 * 4            # rows whose 32 bytes lie inside the index-valued region
            rec = {"layout": name, "start": off, "matches_absolute_address_swizzle": bool((tab == absolute)[rows_ok].all()),
                   "matches_start_relative_swizzle": bool((tab == rel)[rows_ok].all()), "rows_checked": int(rows_ok.sum()),
                   "row0": tab[0].tolist(), "row1": tab[1].tolist(), "row9": tab[9].tolist()}
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tc_decode_kmajor.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
