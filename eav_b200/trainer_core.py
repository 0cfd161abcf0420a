"""Lock-step training of M independent EEGNet models (one per subject) on one GPU.

The reference trains its 42 per-subject models one after the other
(CNN_torch/EEGNet_tor.py:144-162, Dataload_eeg.py:173-256).  They share nothing, so the
B200 design makes "subject" a grid dimension: every kernel launch advances all M models
by one batch (sample n belongs to model n // B), which turns a launch-bound B=32 step
into an M*B-sample step that fills the 148 SMs.  One whole step (index copy, forward,
loss, backward, Adam) is captured in a CUDA graph per (batch size, BN mode) and replayed.

Data stay resident on the device; a step is driven by an int32 index vector (which rows of
the resident dataset form the batch), exactly what the reference's DataLoader produces.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import EAV_DROPOUT_MASK, EAV_DROPOUT_NONE, EAV_DROPOUT_PHILOX
from .ops import EegnetDims, _ptr, _stream


class _Program:
    """Pre-allocated buffers + (optionally) a captured CUDA graph for one (B, bn_train, kind)."""

    def __init__(self, core, B, bn_train, kind):
        self.core, self.B, self.bn_train, self.kind = core, B, bn_train, kind   # kind: 'train' | 'eval'
        dev, M, d = core.device, core.M, core.dims
        n = M * B
        self.idx = torch.zeros(n, dtype=torch.int32, device=dev)
        self.out = torch.empty(n, d.nb_classes, dtype=torch.float32, device=dev)
        self.dout = torch.empty(n, d.nb_classes, dtype=torch.float32, device=dev) if kind == "train" else None
        self.loss = torch.zeros(M, dtype=torch.float32, device=dev)
        self.ncorrect = torch.zeros(M, dtype=torch.int32, device=dev)
        self.mask1 = self.mask2 = None
        self.graph = None
        self.use_x_index = True
        self.x_src = None            # alternative data source (host-fed staging buffer)
        self.y_src = None

    def cfg(self):
        c = self.core
        train_dropout = self.bn_train and c.dims.dropoutRate > 0
        if not train_dropout:
            mode = EAV_DROPOUT_NONE
        elif self.mask1 is not None:
            mode = EAV_DROPOUT_MASK
        else:
            mode = EAV_DROPOUT_PHILOX
        return c.dims.cfg(c.M, self.B, self.bn_train, mode, c.pstride, c.dims.n_bn, seed=c.seed, step=0,
                          step_ptr=c.step_dev.data_ptr() if mode == EAV_DROPOUT_PHILOX else 0)

    def enqueue(self):
        """Issue one step on the current stream (capturable: only kernel launches)."""
        c, lib = self.core, self.core.lib
        cfg = self.cfg()
        x = self.x_src if self.x_src is not None else c.x
        y = self.y_src if self.y_src is not None else c.y
        idx = self.idx if self.use_x_index else None
        st = _stream()
        _lib.check(lib.eav_eegnet_forward(ctypes.byref(cfg), _ptr(x), _ptr(idx), _ptr(c.params), _ptr(c.bn_state),
                                          _ptr(self.mask1), _ptr(self.mask2), _ptr(self.out), _ptr(c.workspace),
                                          c.ws_bytes, st), "eav_eegnet_forward")
        _lib.check(lib.eav_eegnet_loss(ctypes.byref(cfg), _ptr(self.out), _ptr(y), _ptr(idx), _ptr(self.loss),
                                       _ptr(self.dout), _ptr(self.ncorrect), st), "eav_eegnet_loss")
        if self.kind == "train":
            _lib.check(lib.eav_eegnet_backward(ctypes.byref(cfg), _ptr(x), _ptr(idx), _ptr(c.params), _ptr(self.dout),
                                               _ptr(self.mask1), _ptr(self.mask2), _ptr(c.grads), _ptr(c.workspace),
                                               c.ws_bytes, st), "eav_eegnet_backward")
            _lib.check(lib.eav_adam_step_graph(_ptr(c.params), _ptr(c.grads), _ptr(c.exp_avg), _ptr(c.exp_avg_sq),
                                               c.params.numel(), _ptr(c.step_dev), c.lr, c.betas[0], c.betas[1],
                                               c.eps, st), "eav_adam_step_graph")

    def capture(self):
        # warm up on a side stream (sets kernel attributes, allocates nothing), then capture
        step_before = self.core.step_dev.clone()
        snap = self.core.snapshot()
        s = torch.cuda.Stream(device=self.core.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.enqueue()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.enqueue()
        self.graph = g
        # the warm-up really trained one step: roll the state back
        self.core.restore(snap)
        self.core.step_dev.copy_(step_before)
        torch.cuda.synchronize()

    def launch(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.enqueue()


class SubjectBatchTrainer:
    """M models x (params, BN buffers, Adam state) in flat arenas + resident data.

    x: float32 [rows][Chans][Samples] device tensor holding every model's samples,
    y: int64 [rows] labels; a step takes idx int32 [M*B] = absolute row numbers, model-major.
    """

    def __init__(self, dims: EegnetDims, n_models: int, x: torch.Tensor, y: torch.Tensor, lr=1e-4,
                 betas=(0.9, 0.999), eps=1e-8, seed=0, max_batch=32, use_graph=True, params=None,
                 bn_state=None):
        _lib.require_device()
        self.lib = _lib.load()
        self.dims, self.M = dims, n_models
        self.device = x.device
        self.x, self.y = x.contiguous(), y.contiguous()
        if self.x.dtype != torch.float32 or self.y.dtype != torch.int64:
            raise TypeError("x must be float32 and y int64")
        self.lr, self.betas, self.eps, self.seed = float(lr), betas, float(eps), int(seed)
        self.use_graph = use_graph
        self.n_params, self.layout = dims.param_layout()
        dev = self.device
        if params is not None:                                 # adopt the caller's arena (drop-in nn.Module)
            if params.dim() != 2 or params.shape[0] != n_models or params.shape[1] < self.n_params:
                raise ValueError("params arena must be [n_models][>= n_params]")
            self.params = params
        else:
            stride = (self.n_params + 3) // 4 * 4              # 16-byte aligned model slices
            self.params = torch.zeros(n_models, stride, dtype=torch.float32, device=dev)
        self.pstride = self.params.stride(0)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.bn_state = bn_state if bn_state is not None else torch.zeros(n_models, dims.n_bn, dtype=torch.float32, device=dev)
        self._own_bn = bn_state is None
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)   # Adam t (and Philox stream position)
        c = dims.cfg(n_models, max_batch, param_stride=self.pstride, bn_stride=dims.n_bn)
        self.ws_bytes = self.lib.eav_eegnet_workspace_bytes(ctypes.byref(c))
        if self.ws_bytes == 0:
            raise RuntimeError(f"eav_eegnet_workspace_bytes failed: {_lib.last_error()}")
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.max_batch = max_batch
        self._programs = {}
        if self._own_bn:
            self.reset_bn()

    # ----------------------------------------------------------------- state
    def reset_bn(self):
        for i, kind, off, ch in self.dims.bn_layout():
            self.bn_state[:, off:off + ch] = 0.0 if kind == "running_mean" else 1.0

    def load_state_dicts(self, sds, bn_names):
        """sds: list of M reference-named state dicts (CPU tensors)."""
        host = torch.zeros(self.M, self.pstride, dtype=torch.float32)
        bn = torch.zeros(self.M, self.dims.n_bn, dtype=torch.float32)
        for m, sd in enumerate(sds):
            for name, off, shape in self.layout:
                host[m, off:off + int(np.prod(shape))] = sd[name].detach().reshape(-1).float().cpu()
            for i, kind, off, ch in self.dims.bn_layout():
                bn[m, off:off + ch] = sd[f"{bn_names[i]}.{kind}"].detach().float().cpu()
        self.params.copy_(host)
        self.bn_state.copy_(bn)

    def state_dict(self, m, bn_names, num_batches_tracked=0):
        row, bn = self.params[m].detach().cpu(), self.bn_state[m].detach().cpu()
        sd = {name: row[off:off + int(np.prod(shape))].reshape(shape).clone() for name, off, shape in self.layout}
        for i, kind, off, ch in self.dims.bn_layout():
            sd[f"{bn_names[i]}.{kind}"] = bn[off:off + ch].clone()
        for n in bn_names:
            sd[f"{n}.num_batches_tracked"] = torch.tensor(num_batches_tracked)
        return sd

    def snapshot(self):
        return [t.clone() for t in (self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.bn_state)]

    def restore(self, snap):
        for t, s in zip((self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.bn_state), snap):
            t.copy_(s)

    # ----------------------------------------------------------------- steps
    def program(self, B, bn_train, kind="train"):
        key = (B, bool(bn_train), kind)
        p = self._programs.get(key)
        if p is None:
            if B > self.max_batch:
                raise ValueError(f"batch {B} > max_batch {self.max_batch}")
            p = _Program(self, B, bool(bn_train), kind)
            self._programs[key] = p
        return p

    def train_step(self, idx: torch.Tensor, bn_train=True, masks=None):
        """One optimisation step of every model.  idx: int32 device [M*B] absolute rows.
        masks: optional (mask1 [M*B][G][T/4], mask2 [M*B][F2][T/32]) uint8 keep-masks
        (parity mode; not graph-captured).  Returns the per-model loss tensor (device, [M])."""
        B = idx.numel() // self.M
        p = self.program(B, bn_train, "train")
        p.idx.copy_(idx, non_blocking=True)
        if masks is not None:
            p.mask1, p.mask2 = masks
            p.enqueue()
            p.mask1 = p.mask2 = None
            return p.loss
        if self.use_graph and p.graph is None:
            p.capture()
        p.launch()
        return p.loss

    def eval_batch(self, idx: torch.Tensor):
        """Eval-mode forward + loss + #correct for every model on rows idx [M*B]."""
        B = idx.numel() // self.M
        p = self.program(B, False, "eval")
        p.idx.copy_(idx, non_blocking=True)
        if self.use_graph and p.graph is None:
            p.capture()
        p.launch()
        return p.loss, p.ncorrect, p.out

    # -------------------------------------------------------- host-fed steps (end-to-end path)
    def host_step_program(self, B, bn_train=True):
        """A train program whose batch comes from a device staging buffer filled by H2D copies
        (what the reference does every step, EEGNet_tor.py:100-101) instead of resident rows."""
        key = (B, bool(bn_train), "host")
        p = self._programs.get(key)
        if p is None:
            p = _Program(self, B, bool(bn_train), "train")
            d = self.dims
            p.x_src = torch.empty(self.M * B, d.Chans, d.Samples, dtype=torch.float32, device=self.device)
            p.y_src = torch.empty(self.M * B, dtype=torch.int64, device=self.device)
            p.use_x_index = False
            self._programs[key] = p
        return p
