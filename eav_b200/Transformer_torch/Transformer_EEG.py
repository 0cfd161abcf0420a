"""Drop-in for the reference's Transformer_torch/Transformer_EEG.py: `ShallowConvNet`
(Transformer_EEG.py:107-148) and `TrainerUni` (:151-219) with unchanged constructor / forward
signatures, submodule names and state_dict keys, running forward AND backward on hand-written
sm_100a kernels (libeav_b200.so, csrc/shallow.cu) through a torch.autograd.Function, so the
reference's own training loop (`loss = criterion(model(x), y); loss.backward(); optimizer.step()`)
works unmodified.  There is no CPU fallback: forward() on a CPU tensor raises.

The one repair the shipped file needs to run at all is kept: `TrainerUni._loader` is defined without
`self` (Transformer_EEG.py:176) and is called as a bound method, so the reference's constructor raises
TypeError; here it is a static method with the intended behaviour.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader, TensorDataset

from .. import _lib
from .._lib import ShallowCfg
from ..ops import _on, _ptr, _stream


class PatchEmbedding(nn.Module):
    """Transformer_EEG.py:14-34: one Linear(30, 1, bias=False) per temporal filter (parameters only; the
    arithmetic runs inside the fused conv + projection kernel)."""

    def __init__(self, embed_dim: int, num_heads: int, qkv_dim: int):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads, self.qkv_dim = embed_dim, num_heads, qkv_dim
        self.value_proj = nn.ModuleList([nn.Linear(30, 1, bias=False) for _ in range(40)])


class MultiHeadAttention(nn.Module):
    """Transformer_EEG.py:36-69 (parameters only)."""

    def __init__(self, embed_dim: int, num_heads: int, qkv_dim: int):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.W_q = nn.Linear(self.head_dim, qkv_dim, bias=False)
        self.W_k = nn.Linear(self.head_dim, qkv_dim, bias=False)
        self.W_v = nn.Linear(self.head_dim, qkv_dim, bias=False)


class FeedForwardBlock(nn.Module):
    """Transformer_EEG.py:72-84 (parameters only)."""

    def __init__(self, embed_dim: int, expansion: int = 4, drop_p: float = 0.5):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(embed_dim, embed_dim * expansion), nn.ReLU(), nn.Dropout(drop_p),
                                 nn.Linear(embed_dim * expansion, embed_dim))


class TransformerLayer(nn.Module):
    """Transformer_EEG.py:87-103 (parameters only)."""

    def __init__(self, embed_dim: int, num_heads: int, qkv_dim: int, drop_p: float = 0.5):
        super().__init__()
        self.attn = MultiHeadAttention(embed_dim, num_heads, qkv_dim)
        self.ffn = FeedForwardBlock(embed_dim, drop_p=drop_p)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.dropout = nn.Dropout(drop_p)


class _ShallowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x3, masks, flat):
        out = module._run_forward(x3, flat, masks)
        ctx.module, ctx.x3, ctx.masks, ctx.flat, ctx.fwd_id = module, x3, masks, flat, module._fwd_id
        return out

    @staticmethod
    def backward(ctx, dout):
        m = ctx.module
        if m._fwd_id != ctx.fwd_id:
            raise RuntimeError("eav_b200: backward() of a forward whose saved activations were overwritten by a later "
                               "forward (the workspace holds one forward at a time)")
        return None, None, None, m._run_backward(ctx.x3, ctx.flat, ctx.masks, dout.contiguous())


class ShallowConvNet(nn.Module):
    dropout_source = "device"    # 'device': masks drawn on the GPU; 'torch_cpu': from torch's global CPU RNG in the
                                 # reference's call order (parity mode)

    def __init__(self, nb_classes: int, chans: int = 30, samples: int = 500, dropout: float = 0.5, num_layers: int = 12):
        super().__init__()
        # construction order == the reference's (Transformer_EEG.py:118-132): state_dict keys and the default
        # initialisation drawn from torch's global RNG are identical
        self.conv = nn.Conv2d(1, 40, (1, 13), bias=False)
        self.pool = nn.AvgPool2d((1, 35), stride=(1, 7))
        self.dropout = nn.Dropout(dropout)
        self.bn = nn.BatchNorm2d(40)
        self.embedding = PatchEmbedding(embed_dim=40, num_heads=1, qkv_dim=40)
        self.transformer = nn.ModuleList([TransformerLayer(40, 1, 40, dropout) for _ in range(num_layers)])
        self.fc = nn.Linear(2600, nb_classes, bias=False)
        self._nb, self._chans, self._samples, self._p, self._layers = nb_classes, chans, samples, float(dropout), num_layers
        self._ws = {}
        self._fwd_id = 0

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_ws"] = {}
        return st

    # ------------------------------------------------------------------ C ABI plumbing
    def _cfg(self, B, train, with_masks):
        c = ShallowCfg()
        c.batch, c.chans, c.samples, c.n_filters, c.kern = B, self._chans, self._samples, 40, 13
        c.n_layers, c.ffn, c.pool, c.stride, c.n_classes = self._layers, 160, 35, 7, self._nb
        c.bn_train, c.dropout_mode = int(train), int(with_masks)
        c.dropout_p, c.bn_eps, c.bn_momentum, c.ln_eps = self._p, self.bn.eps, self.bn.momentum or 0.1, 1e-5
        return c

    def _flat_params(self):
        """All parameters as one flat fp32 vector in named_parameters() order == the layout of csrc/shallow.cu."""
        return torch.cat([p.reshape(-1) for p in self.parameters()])

    def _workspace(self, B, dev):
        key = (B, dev)
        if key not in self._ws:
            c = self._cfg(B, False, False)
            lib = _lib.load()
            n = lib.eav_shallow_param_layout(ctypes.byref(c), None, None)
            if n != sum(p.numel() for p in self.parameters()):
                raise RuntimeError(f"eav_b200: parameter layout mismatch ({n} vs module)")
            nbytes = lib.eav_shallow_workspace_bytes(ctypes.byref(c))
            if nbytes == 0:
                raise RuntimeError(f"eav_shallow_workspace_bytes failed: {_lib.last_error()}")
            if len(self._ws) >= 3:
                self._ws.clear()
            self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._ws[key]

    def _draw_masks(self, B, dev):
        """uint8 keep flags of every dropout call of one forward, in call order (Transformer_EEG.py:100-101,80,141)."""
        Tp, keep = self._samples - 12, 1.0 - self._p
        shapes = [s for _ in range(self._layers) for s in ((B, Tp, 40), (B, Tp, 160), (B, Tp, 40))] + [(B, 40, (Tp - 35) // 7 + 1)]
        if self.dropout_source == "torch_cpu":
            parts = [torch.empty(s).bernoulli_(keep).to(torch.uint8).reshape(-1) for s in shapes]
            return torch.cat(parts).to(dev)
        n = sum(int(np.prod(s)) for s in shapes)
        return (torch.rand(n, device=dev) < keep).to(torch.uint8)

    def _run_forward(self, x3, flat, masks):
        lib, B, dev = _lib.load(), x3.shape[0], x3.device
        ws = self._workspace(B, dev)
        train = self.training
        c = self._cfg(B, train, masks is not None)
        out = torch.empty(B, self._nb, dtype=torch.float32, device=dev)
        bn = torch.cat([self.bn.running_mean, self.bn.running_var]).contiguous()
        with _on(dev):
            _lib.check(lib.eav_shallow_forward(ctypes.byref(c), _ptr(x3), _ptr(flat), _ptr(bn), _ptr(masks), _ptr(out),
                                               _ptr(ws), ws.numel(), _stream()), "eav_shallow_forward")
        if train:
            with torch.no_grad():
                self.bn.running_mean.copy_(bn[:40]); self.bn.running_var.copy_(bn[40:])
                self.bn.num_batches_tracked += 1
        self._fwd_id += 1
        self._last_cfg = c
        return out

    def _run_backward(self, x3, flat, masks, dout):
        lib, dev = _lib.load(), x3.device
        ws = self._workspace(x3.shape[0], dev)
        grads = torch.empty_like(flat)
        with _on(dev):
            _lib.check(lib.eav_shallow_backward(ctypes.byref(self._last_cfg), _ptr(x3), _ptr(flat), _ptr(dout), _ptr(masks),
                                                _ptr(grads), _ptr(ws), ws.numel(), _stream()), "eav_shallow_backward")
        return grads

    def forward(self, x):
        """x: (B, 1, chans, samples) float32 CUDA -> (B, nb_classes) probabilities (Transformer_EEG.py:122-148)."""
        if not torch.is_tensor(x) or not x.is_cuda:
            raise RuntimeError("eav_b200: forward() needs a CUDA tensor; there is no CPU fallback")
        _lib.require_device()
        if x.dim() != 4 or x.shape[1] != 1 or x.shape[2] != self._chans or x.shape[3] != self._samples:
            raise ValueError(f"expected (B, 1, {self._chans}, {self._samples}), got {tuple(x.shape)}")
        x3 = x.reshape(x.shape[0], self._chans, self._samples).to(torch.float32).contiguous()
        masks = self._draw_masks(x3.shape[0], x3.device) if (self.training and self._p > 0) else None
        flat = self._flat_params()               # differentiable: autograd scatters d(flat) back to every Parameter
        if torch.is_grad_enabled() and flat.requires_grad:
            return _ShallowFn.apply(self, x3, masks, flat)
        return self._run_forward(x3, flat.detach(), masks)


class TrainerUni:
    """Transformer_EEG.py:151-219: same attributes, loop, per-step max-norm on fc.weight (:196-199) and the
    results-file line after the last epoch."""

    def __init__(self, model, data, lr=1e-3, batch_size=32, epochs=10, subject=0, device=None):
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if torch.device(self.device).type != "cuda":
            raise RuntimeError("eav_b200.TrainerUni needs a CUDA (B200) device; there is no CPU fallback")
        tr_x, tr_y, te_x, te_y = data
        self.train_loader = self._loader(tr_x, tr_y, batch_size, True)
        self.test_loader = self._loader(te_x, te_y, batch_size, False)
        self.model = model.to(self.device)
        self.criterion = nn.CrossEntropyLoss()
        self.optimizer = optim.Adam(self.model.parameters(), lr=lr)
        self.epochs = epochs
        self.subject = subject

    @staticmethod
    def _loader(x, y, batch_size, shuffle):
        return DataLoader(TensorDataset(x, y), batch_size=batch_size, shuffle=shuffle)

    def train(self):
        for epoch in range(self.epochs):
            self.model.train()
            for x, y in self.train_loader:
                x, y = x.to(self.device), y.to(self.device)
                out = self.model(x)
                loss = self.criterion(out, y)
                self.optimizer.zero_grad()
                loss.backward()
                self.optimizer.step()
                with torch.no_grad():
                    self.model.fc.weight.data = torch.renorm(self.model.fc.weight.data, p=2, dim=0, maxnorm=0.5)
            acc = self.validate()
            if epoch == self.epochs - 1:
                with open("eeg_results_new_shallow_.txt", "a") as f:
                    f.write(f"Subject {self.subject} | Accuracy: {acc:.4f}\n")

    def validate(self):
        self.model.eval()
        correct, total = 0, 0
        with torch.no_grad():
            for x, y in self.test_loader:
                x, y = x.to(self.device), y.to(self.device)
                preds = self.model(x).argmax(dim=1)
                correct += (preds == y).sum().item()
                total += y.size(0)
        acc = correct / total
        print(f"Validation Accuracy: {acc:.4f}")
        return acc
