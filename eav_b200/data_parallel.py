"""Large-batch single-model training across GPUs (BASELINE.json configs[4]).

One EEGNet replica per rank (one process per GPU), the global batch split evenly.  To equal
the single-device result at the GLOBAL batch size:
  * train-mode BatchNorm statistics are all-reduced (forward: sum / sum of squares of the three
    BN inputs; backward: sum dz / sum dz*xhat) -- 6 tiny float64 buffers per step, none in eval
    mode;
  * the flat fp32 gradient arena (74 933 floats = 300 KB) is all-reduced once per step with
    NCCL over NVLink (latency-bound at this size);
  * every rank then applies the same fused Adam update.
The stages are driven one by one through eav_eegnet_run_stage; eav_eegnet_stage_allreduce
says which workspace region needs a collective after which stage.  The reference's
nn.DataParallel (per-replica BN statistics, hooks on replica weights) is NOT the semantics
reproduced here: this matches the reference run on ONE device at the global batch.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import EAV_DROPOUT_MASK, EAV_DROPOUT_NONE, EAV_DROPOUT_PHILOX
from .ops import EegnetDims, _ptr, _stream


def split_batch(global_batch: int, world: int):
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    return global_batch // world


class DataParallelEEGNet:
    """One model, per-rank batch `B_local`; call step(x_local, y_local) on every rank."""

    def __init__(self, dims: EegnetDims, global_batch: int, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, group=None,
                 device=None, state_dict=None, bn_names=None, seed=0):
        _lib.require_device()
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dims, self.B = dims, split_batch(global_batch, self.world)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_params, self.layout = dims.param_layout()
        self.pstride = (self.n_params + 3) // 4 * 4
        dev = self.device
        self.params = torch.zeros(1, self.pstride, device=dev)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.params), torch.zeros_like(self.params)
        self.bn_state = torch.zeros(1, dims.n_bn, device=dev)
        for i, kind, off, ch in dims.bn_layout():
            self.bn_state[:, off:off + ch] = 0.0 if kind == "running_mean" else 1.0
        self.lr, self.betas, self.eps, self.seed = float(lr), betas, float(eps), int(seed)
        self.t = 0
        c = self._cfg(True, EAV_DROPOUT_NONE)
        self.ws_bytes = self.lib.eav_eegnet_workspace_bytes(ctypes.byref(c))
        if self.ws_bytes == 0:
            raise RuntimeError(_lib.last_error())
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.out = torch.empty(self.B, dims.nb_classes, device=dev)
        self.dout = torch.empty_like(self.out)
        self.loss = torch.zeros(1, device=dev)
        self.ncorrect = torch.zeros(1, dtype=torch.int32, device=dev)
        self.n_stages = self.lib.eav_eegnet_stage_count()
        self.fwd_end = self.lib.eav_eegnet_stage_forward_end()
        if state_dict is not None:
            self.load_state_dict(state_dict, bn_names)

    def _cfg(self, bn_train, mode, step=0):
        return self.dims.cfg(1, self.B, bn_train, mode, self.pstride, self.dims.n_bn, seed=self.seed + self.rank,
                             step=step, dp_world=self.world)

    def load_state_dict(self, sd, bn_names):
        import numpy as np
        host = torch.zeros(1, self.pstride)
        for name, off, shape in self.layout:
            host[0, off:off + int(np.prod(shape))] = sd[name].detach().reshape(-1).float().cpu()
        bn = torch.zeros(1, self.dims.n_bn)
        for i, kind, off, ch in self.dims.bn_layout():
            bn[0, off:off + ch] = sd[f"{bn_names[i]}.{kind}"].detach().float().cpu()
        self.params.copy_(host)
        self.bn_state.copy_(bn)

    def _allreduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def step(self, x, y, bn_train=True, masks=None, update=True):
        """x [B_local][Chans][Samples] f32, y [B_local] i64 (this rank's slice of the global batch).
        Returns the GLOBAL mean loss (device scalar).  masks: optional explicit dropout keep-masks
        for this rank's samples (parity tests); otherwise on-device Philox."""
        d, lib = self.dims, self.lib
        drop = bn_train and d.dropoutRate > 0
        mode = EAV_DROPOUT_NONE if not drop else (EAV_DROPOUT_MASK if masks is not None else EAV_DROPOUT_PHILOX)
        self.t += 1
        cfg = self._cfg(bn_train, mode, step=self.t)
        m1, m2 = masks if masks is not None else (None, None)
        st = _stream()
        off, cnt = ctypes.c_size_t(0), ctypes.c_size_t(0)

        def run(lo, hi):
            for s in range(lo, hi):
                _lib.check(lib.eav_eegnet_run_stage(ctypes.byref(cfg), s, _ptr(x), None, _ptr(self.params),
                                                    _ptr(self.bn_state), _ptr(m1), _ptr(m2), _ptr(self.out),
                                                    _ptr(self.dout), _ptr(self.grads), _ptr(self.workspace),
                                                    self.ws_bytes, st), "eav_eegnet_run_stage")
                _lib.check(lib.eav_eegnet_stage_allreduce(ctypes.byref(cfg), s, ctypes.byref(off), ctypes.byref(cnt)),
                           "eav_eegnet_stage_allreduce")
                if cnt.value:
                    buf = self.workspace[off.value:off.value + 8 * cnt.value].view(torch.float64)
                    self._allreduce(buf)

        run(0, self.fwd_end)
        _lib.check(lib.eav_eegnet_loss(ctypes.byref(cfg), _ptr(self.out), _ptr(y), None, _ptr(self.loss),
                                       _ptr(self.dout), _ptr(self.ncorrect), st), "eav_eegnet_loss")
        self._allreduce(self.loss)
        run(self.fwd_end, self.n_stages)
        self._allreduce(self.grads)                       # one flat 300 KB buffer over NVLink
        if update:
            _lib.check(lib.eav_adam_step(_ptr(self.params), _ptr(self.grads), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                         self.params.numel(), self.t, self.lr, self.betas[0], self.betas[1], self.eps,
                                         st), "eav_adam_step")
        return self.loss[0]
