"""-m gpu tests of the tcgen05 path of the temporal convolution (eav_b200/csrc/tconv_tc.cu):
 * the shared-memory operand address maps the kernels rely on, decoded on the device (eav_tc_probe);
 * the tensor-core kernels against the CUDA-core kernels (EAV_TC=ffma) on the same inputs, for the
   reference shape (EEGNet_tor.py:145: Chans=30, Samples=500, kernLength=300) and for other shapes that
   change the number of k-steps / M tiles.
Parity with the reference itself is in test_gpu_eegnet.py (which runs the tensor-core path by default)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B0 = 64 * 1024
SW128_32B = 1 << 29        # layout_type = 1 in descriptor bits 61..63


def _onehot_image(nfl):
    image = np.zeros(nfl, np.float32)
    image[:2048] = np.arange(2048)             # exact in tf32
    for r in range(8):                         # K-major packed one-hot operand at B0: elem(r, k) = (r == k)
        image[(B0 + (r // 4) * 128 + (r % 8) * 16 + (r % 4) * 4) // 4] = 1.0
    return image


def test_probe_k_major_toeplitz_operand():
    """X[col][i] = xs[4 col + i] through {K-major, no swizzle, LBO 16 B, SBO 128 B} on the raw row."""
    from eav_b200 import ops
    rng = np.random.default_rng(0)
    image = (rng.integers(-8, 9, 24 * 1024) / 4.0).astype(np.float32)
    dev = torch.from_numpy(image).cuda()
    N, ksteps, off = 64, 5, 592
    d, _ = ops.tc_probe(dev, 128, N, ksteps, 1, (off, 16, 128, 0, 32), (B0, 128, 256, 0, N * 32))
    xs = image[off // 4:]
    X = np.stack([xs[4 * c:4 * c + 8 * ksteps] for c in range(128)]).astype(np.float64)
    n, k = np.arange(N)[:, None], np.arange(8 * ksteps)[None, :]
    Bm = image[(B0 + (k // 8) * N * 32 + (n // 8) * 256 + ((k % 8) // 4) * 128 + (n % 8) * 16 + (k % 4) * 4) // 4]
    assert np.array_equal(d.cpu().numpy().astype(np.float64), X @ Bm.astype(np.float64).T)


@pytest.mark.parametrize("as_a", [True, False])
@pytest.mark.parametrize("off,lbo,sbo", [(0, 128, 512), (512, 128, 512), (1024, 2048, 512)])
def test_probe_mn_major_swizzled_operand(as_a, off, lbo, sbo):
    """byte(r, k) = swz(start + (r/32) LBO + (r%32) 4 + (k%4) 128 + (k/4) SBO), swz on the absolute address."""
    from eav_b200 import ops
    dev = torch.from_numpy(_onehot_image(24 * 1024)).cuda()
    onehot = (B0, 128, 256, 0, 0)
    desc = (off, lbo, sbo, 1, 0)
    if as_a:
        d, _ = ops.tc_probe(dev, 128, 32, 1, 1, desc, onehot, a_bits=SW128_32B)
        tab = (d.cpu().numpy()[:, :8] * 4).astype(int)
    else:
        d, _ = ops.tc_probe(dev, 128, 32, 1, 1, onehot, desc, b_bits=SW128_32B)
        tab = (d.cpu().numpy()[:8, :].T * 4).astype(int)
    r, k = np.arange(tab.shape[0])[:, None], np.arange(8)[None, :]
    a = off + (r // 32) * lbo + (r % 32) * 4 + (k % 4) * 128 + (k // 4) * sbo
    assert np.array_equal(tab, a ^ (((a >> 7) & 3) << 5))


def _run(dims, M, B, x, y, params, bn, train, m1, m2, mode):
    from eav_b200.ops import EegnetEngine
    old = os.environ.get("EAV_TC")
    os.environ["EAV_TC"] = mode
    try:
        eng = EegnetEngine(dims, M, B)
        p, b = params.clone(), bn.clone()
        out = eng.forward(x, p, b, bn_train=train, mask1=m1, mask2=m2)
        loss, dout, _ = eng.loss(out, y)
        grads = eng.backward(x, p, dout, mask1=m1, mask2=m2)
        torch.cuda.synchronize()
        return out.clone(), eng.saved("y1").clone(), grads.clone(), b.clone()
    finally:
        if old is None:
            os.environ.pop("EAV_TC", None)
        else:
            os.environ["EAV_TC"] = old


@pytest.mark.parametrize("shape", [
    dict(Chans=30, Samples=500, kernLength=300, M=3, B=32),      # reference shape; 3 models so CTAs cross models
    dict(Chans=30, Samples=500, kernLength=300, M=1, B=5),       # short row ranges
    dict(Chans=8, Samples=256, kernLength=64, M=2, B=16),        # CNN_EEG.py:12 style kernLength, 1 k-step of dW
    dict(Chans=30, Samples=512, kernLength=301, M=1, B=8),       # odd kernel, T = 512 (all 128 columns valid)
    dict(Chans=4, Samples=128, kernLength=16, M=2, B=8),         # tiny
])
@pytest.mark.parametrize("train", [True, False])
def test_tensor_core_path_matches_cuda_core_path(shape, train):
    from eav_b200.ops import EegnetDims
    import gpu_util as U
    M, B = shape["M"], shape["B"]
    dims = EegnetDims(5, Chans=shape["Chans"], Samples=shape["Samples"], kernLength=shape["kernLength"])
    n_params, _ = dims.param_layout()
    gen = torch.Generator().manual_seed(7)
    params = (torch.randn(M, n_params, generator=gen) * 0.1).cuda()
    bn = torch.zeros(M, dims.n_bn)
    for i, kind, off, ch in dims.bn_layout():
        bn[:, off:off + ch] = 0.1 * torch.randn(M, ch, generator=gen) if kind == "running_mean" else \
            1.0 + 0.2 * torch.rand(M, ch, generator=gen)
    bn = bn.cuda()
    N = M * B
    x = torch.randn(N, dims.Chans, dims.Samples, generator=gen).cuda()
    y = torch.randint(0, 5, (N,), generator=gen).cuda()
    T4 = dims.Samples // 4
    m1 = (torch.rand(N, 64, T4, generator=gen) > 0.5).to(torch.uint8).cuda() if train else None
    m2 = (torch.rand(N, 64, T4 // 8, generator=gen) > 0.5).to(torch.uint8).cuda() if train else None
    ref = _run(dims, M, B, x, y, params, bn, train, m1, m2, "ffma")
    got = _run(dims, M, B, x, y, params, bn, train, m1, m2, "tc")
    names = ("out", "y1", "grads", "bn_state")
    for nm, a, b in zip(names, got, ref):
        assert torch.isfinite(a).all(), nm
        err = U.rel_l2(a.cpu().numpy(), b.cpu().numpy())
        assert err < 2e-5, (nm, err)
    # the temporal-conv weight gradient on its own (it is a small part of the gradient arena)
    _, layout = dims.param_layout()
    name, off, shp = layout[0]
    n = int(np.prod(shp))
    err = U.rel_l2(got[2][:, off:off + n].cpu().numpy(), ref[2][:, off:off + n].cpu().numpy())
    assert err < 2e-5, (name, err)
