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


def bench_subject_inputs():
    """One subject of the BENCHMARK configuration (bench.py / BASELINE configs[0,2]): 280 train + 120 test epochs of
    30 x 500, labels 0..4.  A class-dependent 10 Hz component gives the model something to learn."""
    g = torch.Generator().manual_seed(14)
    w = torch.randn(5, 30, generator=g)
    osc = torch.sin(torch.arange(500) * (2 * np.pi * 10.0 / 100.0))

    def make(n):
        y = torch.randint(0, 5, (n,), generator=g)
        x = torch.randn(n, 1, 30, 500, generator=g) + 0.5 * (w[y].reshape(n, 1, 30, 1) * osc)
        return x.numpy(), y.numpy()
    trx, try_ = make(280)
    tex, tey = make(120)
    return trx, try_, tex, tey


def cnn_trainer_inputs():
    """EEGNetTrainer golden (CNN_EEG defaults Chans=64, Samples=128): 88 train / 40 validation samples, 4 classes."""
    g = torch.Generator().manual_seed(15)
    trx, tex = torch.randn(88, 64, 128, generator=g), torch.randn(40, 64, 128, generator=g)
    try_, tey = torch.randint(0, 4, (88,), generator=g), torch.randint(0, 4, (40,), generator=g)
    return trx, try_, tex, tey
