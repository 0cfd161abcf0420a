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
from .ops import EegnetDims, _on, _ptr, _stream


class PeerExchange:
    """In-place sum all-reduce of small CUDA buffers by ONE kernel per rank that reads the peers' memory over
    NVLink (eav_peer_allreduce) instead of an NCCL call.  The exchange buffer is torch symmetric memory (mapped
    into every rank of the group); every rank must call allreduce() in the same order."""

    def __init__(self, slot_bytes, group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        nbytes = int(self.lib.eav_peer_exchange_bytes(self.slot_bytes))
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        torch.cuda.synchronize(dev)
        self.handle = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.bases = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=dev)
        dist.barrier(group=self.group, device_ids=[dev.index])      # every rank zeroed its flags before the first call
        self.call = 0

    def allreduce(self, t, call=None, call_base=None, calls_per_step=0):
        """Eager numbering: call=None (an internal host counter).  Graph-safe numbering: call in 1..calls_per_step
        and call_base = int64 device tensor holding the step counter (effective number = step * calls_per_step + call)."""
        if not (t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64)):
            raise ValueError("PeerExchange.allreduce takes a contiguous float32 / float64 CUDA tensor")
        if call is None:
            self.call += 1
            call = self.call & 0xFFFFFFFF or 1
        with _on(t.device):
            _lib.check(self.lib.eav_peer_allreduce(_ptr(t), _ptr(t), t.numel(), int(t.dtype == torch.float64),
                                                   _ptr(self.bases), self.world, self.rank, self.slot_bytes, int(call),
                                                   _ptr(call_base), int(calls_per_step), _stream()), "eav_peer_allreduce")
        return t


def split_batch(global_batch: int, world: int):
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    return global_batch // world


class DataParallelEEGNet:
    """One model, per-rank batch `B_local`; call step(x_local, y_local) on every rank."""

    def __init__(self, dims: EegnetDims, global_batch: int, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, group=None,
                 device=None, state_dict=None, bn_names=None, seed=0, collective="auto"):
        """collective: "peer" = eav_peer_allreduce over symmetric memory (one kernel per reduction, bit-identical
        results on every rank), "nccl" = torch.distributed.all_reduce, "auto" = peer when the symmetric-memory
        rendezvous works, else nccl."""
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
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)    # completed steps: Adam t - 1, Philox position, peer call base
        self._graphs = {}
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
        self.peer = None
        if self.world > 1 and collective in ("auto", "peer"):
            try:
                self.peer = PeerExchange(4 * self.pstride, group=group, device=dev)
            except Exception:
                if collective == "peer":
                    raise
        self.collective = "peer" if self.peer is not None else ("nccl" if self.world > 1 else "none")

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

    CALLS_PER_STEP = 8     # 3 BN forward + loss + 3 BN backward + gradient arena

    def _allreduce(self, t):
        if self.peer is not None:
            self._k += 1
            self.peer.allreduce(t, call=self._k, call_base=self.step_dev, calls_per_step=self.CALLS_PER_STEP)
        elif self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _enqueue(self, x, y, bn_train, mode, m1, m2, update):
        """One whole step on the current stream, no host synchronisation: capturable in a CUDA graph when the
        collective is "peer" (or world == 1).  The step number lives on the device (self.step_dev)."""
        lib = self.lib
        self._k = 0
        cfg = self.dims.cfg(1, self.B, bn_train, mode, self.pstride, self.dims.n_bn, seed=self.seed + self.rank,
                            step=0, step_ptr=self.step_dev.data_ptr(), dp_world=self.world)
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
        if update:     # t = step_dev + 1 read on the device; the trailing kernel increments step_dev
            _lib.check(lib.eav_adam_step_graph(_ptr(self.params), _ptr(self.grads), _ptr(self.exp_avg),
                                               _ptr(self.exp_avg_sq), self.params.numel(), _ptr(self.step_dev),
                                               self.lr, self.betas[0], self.betas[1], self.eps, st), "eav_adam_step_graph")
        else:
            self.step_dev += 1

    def step(self, x, y, bn_train=True, masks=None, update=True, graph=False):
        with _on(self.device):
            return self._step(x, y, bn_train, masks, update, graph)

    def _step(self, x, y, bn_train=True, masks=None, update=True, graph=False):
        """x [B_local][Chans][Samples] f32, y [B_local] i64 (this rank's slice of the global batch).
        Returns the GLOBAL mean loss (device scalar).  masks: optional explicit dropout keep-masks
        for this rank's samples (parity tests); otherwise on-device Philox.
        graph=True replays a captured CUDA graph of the whole step (inputs are copied into static buffers);
        it needs the "peer" collective (or a single rank) and on-device dropout."""
        d = self.dims
        drop = bn_train and d.dropoutRate > 0
        mode = EAV_DROPOUT_NONE if not drop else (EAV_DROPOUT_MASK if masks is not None else d.philox_mode)
        m1, m2 = masks if masks is not None else (None, None)
        self.t += 1
        if not graph:
            self._enqueue(x, y, bn_train, mode, m1, m2, update)
            return self.loss[0]
        if masks is not None or (self.world > 1 and self.peer is None):
            raise ValueError("graph=True needs on-device dropout and the peer collective")
        key = (bool(bn_train), bool(update))
        if key not in self._graphs:
            xs, ys = torch.empty_like(x), torch.empty_like(y)
            xs.copy_(x); ys.copy_(y)
            self._enqueue(xs, ys, bn_train, mode, None, None, update)      # eager once: one-time attribute setup
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):       # recording only: the warm-up above WAS this call's step
                self._enqueue(xs, ys, bn_train, mode, None, None, update)
            self._graphs[key] = (g, xs, ys)
            return self.loss[0]
        g, xs, ys = self._graphs[key]
        xs.copy_(x); ys.copy_(y)
        g.replay()
        return self.loss[0]
