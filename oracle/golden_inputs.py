"""Seeded inputs shared by oracle/gen_golden.py (which feeds them to the reference)
and the tests (which feed them to the oracle / CUDA path).  Large inputs are
regenerated from their seed instead of being stored; each fixture carries a
float64 checksum so a drifting RNG implementation fails loudly.
TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch


def checksum(*arrs):
    return np.array([float(np.asarray(a, dtype=np.float64).sum()) for a in arrs]
                    + [float(np.abs(np.asarray(a, dtype=np.float64)).sum()) for a in arrs])


def adam6_inputs():
    g = torch.Generator().manual_seed(12)
    xs = torch.randn(6, 8, 1, 30, 500, generator=g)
    ys = torch.randint(0, 5, (6, 8), generator=g)
    return xs, ys


def trainer_inputs():
    g = torch.Generator().manual_seed(13)
    trx, tex = torch.randn(40, 1, 30, 500, generator=g).numpy(), torch.randn(16, 1, 30, 500, generator=g).numpy()
    try_, tey = torch.randint(0, 5, (40,), generator=g).numpy(), torch.randint(0, 5, (16,), generator=g).numpy()
    return trx, try_, tex, tey
