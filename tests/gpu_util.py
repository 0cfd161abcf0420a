"""Helpers for the -m gpu parity tests: pack reference-named state dicts into the flat
arenas the C ABI takes, and recompute layer-by-layer activations with the CPU oracle."""
import numpy as np
import torch
import torch.nn.functional as F

from eav_b200.ops import EegnetDims, EegnetEngine
from eav_b200._lib import EAV_VARIANT_CNN, EAV_VARIANT_TOR

TOR_BN = ("firstBN", "depthwiseBN", "separableBN")
CNN_BN = ("block1.1", "block1.3", "block2.2")


def init_from_golden(g):
    return {k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("init::")}


def pack_params(dims, sds, device="cuda"):
    n, layout = dims.param_layout()
    arena = torch.zeros(len(sds), n, dtype=torch.float32)
    for m, sd in enumerate(sds):
        for name, off, shape in layout:
            arena[m, off:off + int(np.prod(shape))] = sd[name].reshape(-1).float()
    return arena.to(device)


def pack_bn(dims, sds, device="cuda"):
    bn_names = TOR_BN if dims.variant == EAV_VARIANT_TOR else CNN_BN
    arena = torch.zeros(len(sds), dims.n_bn, dtype=torch.float32)
    for m, sd in enumerate(sds):
        for i, kind, off, ch in dims.bn_layout():
            arena[m, off:off + ch] = sd[f"{bn_names[i]}.{kind}"].float()
    return arena.to(device)


def unpack(dims, arena_row):
    """arena row (P,) -> {name: tensor(shape)} on CPU."""
    _, layout = dims.param_layout()
    row = arena_row.detach().cpu()
    return {name: row[off:off + int(np.prod(shape))].reshape(shape).clone() for name, off, shape in layout}


def unpack_bn(dims, bn_row):
    bn_names = TOR_BN if dims.variant == EAV_VARIANT_TOR else CNN_BN
    row = bn_row.detach().cpu()
    return {f"{bn_names[i]}.{kind}": row[off:off + ch].clone() for i, kind, off, ch in dims.bn_layout()}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a - b)


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    m = np.abs(b).max()
    return np.abs(a - b).max() / m if m > 0 else np.abs(a - b).max()


def oracle_intermediates_tor(sd, x, train, masks=None, p=0.5):
    """Layer-by-layer activations of EEGNet_tor on CPU (fp32), named like the workspace."""
    out = {}
    W1, W2, W3 = sd["firstConv.weight"], sd["depthwiseConv.weight"], sd["separableConv.weight"]
    bn = lambda h, pre: F.batch_norm(h, sd[pre + ".running_mean"].clone(), sd[pre + ".running_var"].clone(),
                                     sd[pre + ".weight"], sd[pre + ".bias"], training=train, momentum=0.1, eps=1e-5)
    k = W1.shape[-1]
    h = F.conv2d(F.pad(x, ((k - 1) // 2, k - 1 - (k - 1) // 2)), W1)
    out["y1"] = h
    h = F.elu(bn(h, "firstBN"))
    h = F.conv2d(h, W2, groups=W1.shape[0])
    out["y2"] = h.squeeze(2)
    h = F.avg_pool2d(F.elu(bn(h, "depthwiseBN")), (1, 4))
    if train and masks is not None:
        h = h * masks[0].float() / (1 - p)
    out["d1"] = h.squeeze(2)
    k = W3.shape[-1]
    h = F.conv2d(F.pad(h, ((k - 1) // 2, k - 1 - (k - 1) // 2)), W3)
    out["y3"] = h.squeeze(2)
    h = F.avg_pool2d(F.elu(bn(h, "separableBN")), (1, 8))
    if train and masks is not None:
        h = h * masks[1].float() / (1 - p)
    out["feat"] = h.flatten(1)
    return out
