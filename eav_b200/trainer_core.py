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
import os

import numpy as np
import torch

from . import _lib
from ._lib import EAV_DROPOUT_MASK, EAV_DROPOUT_NONE, EAV_DROPOUT_PHILOX
from .ops import EegnetDims, _on, _ptr, _stream


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
        self.params = self.bn_state = self.workspace = None   # optional overrides (validation on a parameter snapshot)

    def cfg(self):
        c = self.core
        train_dropout = self.bn_train and c.dims.dropoutRate > 0
        if not train_dropout:
            mode = EAV_DROPOUT_NONE
        elif self.mask1 is not None:
            mode = EAV_DROPOUT_MASK
        else:
            mode = c.dims.philox_mode
        return c.dims.cfg(c.M, self.B, self.bn_train, mode, c.pstride, c.dims.n_bn, seed=c.seed, step=0,
                          step_ptr=c.step_dev.data_ptr() if mode >= EAV_DROPOUT_PHILOX else 0)

    def enqueue(self):
        """Issue one step on the current stream of the trainer's device (capturable: only kernel launches)."""
        with _on(self.core.device):
            self._enqueue()

    def _enqueue(self):
        c, lib = self.core, self.core.lib
        cfg = self.cfg()
        x = self.x_src if self.x_src is not None else c.x
        y = self.y_src if self.y_src is not None else c.y
        idx = self.idx if self.use_x_index else None
        st = _stream()
        params = self.params if self.params is not None else c.params
        bn_state = self.bn_state if self.bn_state is not None else c.bn_state
        workspace = self.workspace if self.workspace is not None else c.workspace
        _lib.check(lib.eav_eegnet_forward(ctypes.byref(cfg), _ptr(x), _ptr(idx), _ptr(params), _ptr(bn_state),
                                          _ptr(self.mask1), _ptr(self.mask2), _ptr(self.out), _ptr(workspace),
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
        with _on(self.core.device):
            self._capture()

    def _capture(self):
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

    def run(self):
        """launch(), capturing the CUDA graph on first use when the trainer uses graphs."""
        if self.graph is None and self.core.use_graph and self.mask1 is None:
            self.capture()
        self.launch()

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

    def set_lr(self, lr):
        """New learning rate for every later step (the value is a launch argument baked into captured graphs, so the
        train programs are re-captured on their next use)."""
        lr = float(lr)
        if lr == self.lr:
            return
        self.lr = lr
        for p in self._programs.values():
            if p.kind == "train":
                p.graph = None

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

    def epoch_runner(self, n_train, n_test, batch, rows_per_model=None, train_first_row=0, test_first_row=None,
                     subject_ids=None, max_epochs=1024, seed=None, pipeline_validation=True):
        """Whole-epoch CUDA graphs over this trainer's resident rows (see EpochRunner)."""
        return EpochRunner(self, n_train, n_test, batch, rows_per_model, train_first_row, test_first_row,
                           subject_ids, max_epochs, seed, pipeline_validation)

    # -------------------------------------------------------- host-fed steps (end-to-end path)
    def host_step_program(self, B, bn_train=True, kind="train", x_src=None, y_src=None, slot=0):
        """A program whose batch comes from a device staging buffer filled by H2D copies (what the
        reference does every step, EEGNet_tor.py:100-101, and every validation batch, :124-125) instead
        of resident rows.  x_src / y_src: caller-owned staging buffers (>= M*B rows); `slot`
        distinguishes programs bound to different buffers of a double-buffered staging area."""
        key = (B, bool(bn_train), "host", kind, slot)
        p = self._programs.get(key)
        if p is None:
            p = _Program(self, B, bool(bn_train), kind)
            d = self.dims
            p.x_src = x_src if x_src is not None else torch.empty(self.M * B, d.Chans, d.Samples, dtype=torch.float32,
                                                                  device=self.device)
            p.y_src = y_src if y_src is not None else torch.empty(self.M * B, dtype=torch.int64, device=self.device)
            p.use_x_index = False
            self._programs[key] = p
        return p

class EpochRunner:
    """One CUDA graph = one epoch of Trainer_uni.train() (EEGNet_tor.py:96-116) for all M models:
    a fresh on-device permutation (eav_epoch_schedule), ceil(n_train/batch) train steps including
    the ragged last batch (drop_last=False, EEGNet_tor.py:92-93), the validation pass over the
    n_test rows (EEGNet_tor.py:118-135) and the per-epoch loss / accuracy bookkeeping -- no host
    work, no H2D copy and no synchronisation between steps.  The host replays one graph per epoch
    and reads the [epochs][M][3] history (mean train loss, mean validation loss, validation
    accuracy) once at the end.

    Row layout of core.x / core.y: model m's training rows are train_first_row + m*rows_per_model
    + [0, n_train), its test rows test_first_row + m*rows_per_model + [0, n_test).
    Graphs: train-mode BN + dropout (the reference's epoch 1) and eval-mode BN (every later
    epoch), each with and without a validation branch.

    pipeline_validation (default): the validation pass of epoch e runs on a SNAPSHOT of the parameters
    taken at the end of epoch e, as a parallel branch of the graph that trains epoch e+1 -- the same
    numbers as validating in between (the snapshot is what validate() would have seen), but the forward-only
    validation kernels fill the SMs the small training kernels leave idle (5-6 models per GPU when the 42
    subjects are spread over 8 GPUs).  results() / finish() run the last epoch's validation.
    """

    def __init__(self, core, n_train, n_test, batch, rows_per_model=None, train_first_row=0, test_first_row=None,
                 subject_ids=None, max_epochs=1024, seed=None, pipeline_validation=True):
        self.core, self.n_train, self.n_test, self.batch = core, int(n_train), int(n_test), int(batch)
        self.rows_per_model = int(rows_per_model if rows_per_model is not None else n_train + n_test)
        self.train_first_row = int(train_first_row)
        self.test_first_row = int(test_first_row if test_first_row is not None else train_first_row + n_train)
        self.max_epochs = int(max_epochs)
        self.seed = int(core.seed if seed is None else seed) & (2 ** 64 - 1)
        if batch > core.max_batch:
            raise ValueError(f"batch {batch} > max_batch {core.max_batch}")
        dev, M = core.device, core.M
        self.train_sizes = [min(batch, n_train - b0) for b0 in range(0, n_train, batch)]
        self.val_sizes = [min(batch, n_test - b0) for b0 in range(0, n_test, batch)]
        self.sched = torch.zeros(max(1, len(self.train_sizes)), M * batch, dtype=torch.int32, device=dev)
        ids = torch.arange(M) if subject_ids is None else torch.as_tensor(list(subject_ids))
        if ids.numel() != M:
            raise ValueError("subject_ids must have one entry per model")
        self.subject_ids = ids.to(torch.int32).to(dev)
        base = (torch.arange(M) * self.rows_per_model + self.test_first_row).unsqueeze(1)
        self.val_idx = []
        for v, Bv in enumerate(self.val_sizes):
            cols = torch.arange(v * batch, v * batch + Bv).unsqueeze(0)
            self.val_idx.append((base + cols).reshape(-1).to(torch.int32).to(dev))
        self.epoch_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.train_acc = torch.zeros(M, 2, dtype=torch.float64, device=dev)
        self.val_acc = torch.zeros(M, 2, dtype=torch.float64, device=dev)
        self.history = torch.zeros(self.max_epochs, M, 3, dtype=torch.float32, device=dev)
        self._graphs = {}
        if os.environ.get("EAV_PIPELINE_VAL") in ("0", "1"):           # A/B switch for measurements
            pipeline_validation = os.environ["EAV_PIPELINE_VAL"] == "1"
        self.pipeline = bool(pipeline_validation) and len(self.val_sizes) > 0
        self._val_pending = False
        if self.pipeline:
            self.params_val, self.bn_val = core.params.clone(), core.bn_state.clone()
            self.ws_val = torch.empty(core.ws_bytes, dtype=torch.uint8, device=dev)
            self._val_stream = torch.cuda.Stream(device=dev)

    # ------------------------------------------------------------------------------
    def step_index(self, s):
        """x_index vector of train step s of the CURRENT schedule buffer (device view, [M*B_s])."""
        M, B = self.core.M, self.batch
        return self.sched[s, :M * self.train_sizes[s]]

    def enqueue_schedule(self, sched=None, epoch_dev=None):
        c = self.core
        sched = self.sched if sched is None else sched
        epoch_dev = self.epoch_dev if epoch_dev is None else epoch_dev
        with _on(c.device):
            self._enqueue_schedule(sched, epoch_dev)

    def _enqueue_schedule(self, sched, epoch_dev):
        c = self.core
        _lib.check(c.lib.eav_epoch_schedule(_ptr(sched), _ptr(self.subject_ids), c.M, self.n_train, self.batch,
                                            self.rows_per_model, self.train_first_row, self.seed, _ptr(epoch_dev),
                                            _stream()), "eav_epoch_schedule")

    def peek_schedule(self, epoch):
        """Host copy of the index schedule of `epoch` (list of int32 CPU tensors, one per step)."""
        tmp = torch.zeros_like(self.sched)
        e = torch.tensor([int(epoch)], dtype=torch.int64, device=self.core.device)
        self.enqueue_schedule(tmp, e)
        M = self.core.M
        return [tmp[s, :M * Bs].cpu() for s, Bs in enumerate(self.train_sizes)]

    def enqueue(self, bn_train, with_val=False):
        """Issue one whole epoch on the current stream (kernel launches only: capturable)."""
        with _on(self.core.device):
            self._enqueue(bn_train, with_val)

    def _enqueue_train(self, bn_train):
        c, lib = self.core, self.core.lib
        self.enqueue_schedule()
        for s, Bs in enumerate(self.train_sizes):
            p = c.program(Bs, bn_train, "train")
            keep = p.idx
            p.idx = self.step_index(s)
            try:
                p.enqueue()
            finally:
                p.idx = keep
            _lib.check(lib.eav_epoch_accumulate(_ptr(p.loss), None, c.M, _ptr(self.train_acc), _stream()),
                       "eav_epoch_accumulate")

    def _enqueue_val(self, snapshot):
        """The validation pass (EEGNet_tor.py:118-135) on the live parameters, or on the snapshot with its own workspace."""
        c, lib = self.core, self.core.lib
        for v, Bv in enumerate(self.val_sizes):
            p = c.program(Bv, False, "eval_snap" if snapshot else "eval")
            keep = p.idx
            p.idx = self.val_idx[v]
            if snapshot:
                p.params, p.bn_state, p.workspace = self.params_val, self.bn_val, self.ws_val
            try:
                p.enqueue()
            finally:
                p.idx = keep
            _lib.check(lib.eav_epoch_accumulate(_ptr(p.loss), _ptr(p.ncorrect), c.M, _ptr(self.val_acc), _stream()),
                       "eav_epoch_accumulate")

    def _commit(self, with_val_of_this_epoch):
        c, lib = self.core, self.core.lib
        _lib.check(lib.eav_epoch_commit(_ptr(self.train_acc), _ptr(self.val_acc) if with_val_of_this_epoch else None, c.M,
                                        len(self.train_sizes), len(self.val_sizes), self.n_test, _ptr(self.history),
                                        self.max_epochs, _ptr(self.epoch_dev), _stream()), "eav_epoch_commit")

    def _commit_val_of_previous_epoch(self):
        c, lib = self.core, self.core.lib
        _lib.check(lib.eav_epoch_commit_val(_ptr(self.val_acc), c.M, len(self.val_sizes), self.n_test, _ptr(self.history),
                                            self.max_epochs, _ptr(self.epoch_dev), -1, _stream()), "eav_epoch_commit_val")

    def _enqueue(self, bn_train, with_val=False):
        c = self.core
        if not self.pipeline:
            self._enqueue_train(bn_train)
            self._enqueue_val(False)
            self._commit(True)
            return
        cur = torch.cuda.current_stream()
        if with_val:                                    # previous epoch's validation: a parallel branch on the snapshot
            self._val_stream.wait_stream(cur)
            with torch.cuda.stream(self._val_stream):
                self._enqueue_val(True)
        self._enqueue_train(bn_train)
        if with_val:
            cur.wait_stream(self._val_stream)
            self._commit_val_of_previous_epoch()
        self.params_val.copy_(c.params)                 # what validate() sees at the end of this epoch
        self.bn_val.copy_(c.bn_state)
        # validate()'s forward also leaves its max-norm hooks' renorm in the LIVE weights (EEGNet_tor.py:33-34,47-48);
        # the pipelined pass will only renorm the snapshot
        cfg = c.dims.cfg(c.M, 1, param_stride=c.pstride, bn_stride=c.dims.n_bn)
        _lib.check(c.lib.eav_eegnet_apply_hooks(ctypes.byref(cfg), _ptr(c.params), _stream()), "eav_eegnet_apply_hooks")
        self._commit(False)

    def _state(self):
        c = self.core
        st = (c.params, c.grads, c.exp_avg, c.exp_avg_sq, c.bn_state, c.step_dev, self.epoch_dev, self.train_acc,
              self.val_acc, self.history)
        return st + ((self.params_val, self.bn_val) if self.pipeline else ())

    def capture(self, bn_train, with_val=False):
        """Warm-up (really runs one epoch on a side stream), capture, roll every piece of state back."""
        with _on(self.core.device):
            return self._capture(bn_train, with_val)

    def _capture(self, bn_train, with_val):
        snap = [t.clone() for t in self._state()]
        s = torch.cuda.Stream(device=self.core.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.enqueue(bn_train, with_val)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.enqueue(bn_train, with_val)
        for t, v in zip(self._state(), snap):
            t.copy_(v)
        torch.cuda.synchronize()
        self._graphs[(bool(bn_train), bool(with_val))] = g
        return g

    def run_epoch(self, bn_train, use_graph=True):
        """Advance every model by one epoch (asynchronous: returns after the launch)."""
        with_val = self.pipeline and self._val_pending
        if not use_graph:
            self.enqueue(bn_train, with_val)
        else:
            g = self._graphs.get((bool(bn_train), bool(with_val)))
            if g is None:
                g = self.capture(bn_train, with_val)
            g.replay()
        self._val_pending = self.pipeline

    def finish(self):
        """Pipelined mode: the validation pass of the last trained epoch (no-op otherwise / when none is pending)."""
        if not (self.pipeline and self._val_pending):
            return
        with _on(self.core.device):
            self._enqueue_val(True)
            self._commit_val_of_previous_epoch()
        self._val_pending = False

    def epochs_done(self):
        return int(self.epoch_dev.item())

    def results(self):
        """(history [epochs_done][M][3] CPU float32) -- one synchronisation."""
        self.finish()
        n = self.epochs_done()
        return self.history[:min(n, self.max_epochs)].cpu()

    @property
    def samples_per_epoch(self):
        return self.core.M * self.n_train
