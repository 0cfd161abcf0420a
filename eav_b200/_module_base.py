"""Shared machinery of the drop-in nn.Modules and trainers.

ArenaModule keeps every parameter of the reference-named submodules as a VIEW into one flat
device arena (the layout libeav_b200.so reads), so `state_dict()`, `optimizer.step()`,
`load_state_dict()` and the CUDA kernels all see the same memory.  forward() goes through a
torch.autograd.Function that calls the C ABI, so unmodified trainer code
(`loss = criterion(model(x), y); loss.backward(); optimizer.step()`) runs on the kernels.

FusedTrainerMixin drives the fully fused path (resident dataset, index-driven batches,
forward + loss + backward + Adam captured in one CUDA graph) for the drop-in trainers.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader, TensorDataset

from . import _lib
from .ops import EegnetEngine
from .trainer_core import SubjectBatchTrainer


class _EEGNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x3, masks, philox, *params):
        eng = module._engine(x3.shape[0])
        train = module.training
        out = eng.forward(x3, module._arena, module._bn_arena, bn_train=train,
                          mask1=masks[0] if masks else None, mask2=masks[1] if masks else None, philox=philox)
        module._fwd_id += 1
        ctx.module, ctx.eng, ctx.x3, ctx.masks, ctx.fwd_id = module, eng, x3, masks, module._fwd_id
        eng._owner_fwd_id = module._fwd_id
        return out

    @staticmethod
    def backward(ctx, dout):
        m, eng = ctx.module, ctx.eng
        if eng._owner_fwd_id != ctx.fwd_id:
            raise RuntimeError("eav_b200: backward() of a forward whose saved activations were overwritten by a "
                               "later forward of the same batch size (the workspace holds one forward at a time)")
        grads = eng.backward(ctx.x3, m._arena, dout.contiguous(), mask1=ctx.masks[0] if ctx.masks else None,
                             mask2=ctx.masks[1] if ctx.masks else None)
        outs = []
        for name, off, shape in eng.layout:
            outs.append(grads[0, off:off + int(np.prod(shape))].view(shape).clone())
        return (None, None, None, None) + tuple(outs)


class ArenaModule(nn.Module):
    """Base of EEGNet_tor / EEGNet: subclasses set _dims, _param_modules, _BN_NAMES, _dropout2d."""

    dropout_source = "philox"   # 'philox': on-device dropout; 'torch_cpu': masks from torch's global CPU
                                # RNG in the reference's draw order (parity mode)

    def __init__(self):
        super().__init__()
        self._arena = None
        self._bn_arena = None
        self._engines = {}
        self._fwd_id = 0
        self._philox_step = 0

    # ------------------------------------------------------------ copy / pickle
    def __getstate__(self):
        """copy.deepcopy(model) (best-model snapshots), torch.save(model) and pickling work after a forward like for
        the reference nn.Module: the engines (ctypes handles, workspaces) and the arena handles are not part of the
        state -- the copy's parameters stay views of ONE copied storage and _ensure_arena() re-adopts or repacks them
        on its next forward."""
        st = self.__dict__.copy()
        st["_engines"] = {}
        st["_arena"] = None
        st["_bn_arena"] = None
        return st

    # ------------------------------------------------------------ arena management
    def _named_param_list(self):
        mods = dict(self.named_parameters())
        return [mods[n] for n in self._param_modules]

    def _bn_buffers(self):
        out = []
        for i, kind, off, ch in self._dims.bn_layout():
            out.append((getattr(self.get_submodule(self._BN_NAMES[i]), kind), off, ch))
        return out

    def _arena_ok(self):
        if self._arena is None or not self._arena.is_cuda:
            return False
        base = self._arena.data_ptr()
        for p, (name, off, shape) in zip(self._named_param_list(), self._layout):
            if p.device != self._arena.device or p.data_ptr() != base + 4 * off or not p.is_contiguous():
                return False
        bbase = self._bn_arena.data_ptr()
        for buf, off, ch in self._bn_buffers():
            if buf.device != self._bn_arena.device or buf.data_ptr() != bbase + 4 * off:
                return False
        return True

    def _ensure_arena(self):
        """(Re)pack parameters and BN running statistics into the flat arenas and re-point
        every Parameter / buffer at its slice.  Cheap pointer check when already packed."""
        if not hasattr(self, "_layout"):
            self._n_params, self._layout = self._dims.param_layout()
        if self._arena_ok():
            return
        plist = self._named_param_list()
        dev = plist[0].device
        if dev.type != "cuda":
            raise RuntimeError("eav_b200: the model must be on a CUDA (B200) device; there is no CPU fallback")
        stride = (self._n_params + 3) // 4 * 4
        arena = torch.zeros(1, stride, dtype=torch.float32, device=dev)
        for p, (name, off, shape) in zip(plist, self._layout):
            n = int(np.prod(shape))
            arena[0, off:off + n] = p.detach().reshape(-1).to(dev, torch.float32)
            p.data = arena[0, off:off + n].view(shape)
        bn = torch.zeros(1, self._dims.n_bn, dtype=torch.float32, device=dev)
        for buf, off, ch in self._bn_buffers():
            bn[0, off:off + ch] = buf.detach().to(dev, torch.float32)
            buf.data = bn[0, off:off + ch]
        self._arena, self._bn_arena = arena, bn
        self._engines = {}

    def _engine(self, B):
        eng = self._engines.get(B)
        if eng is None:
            if len(self._engines) >= 6:
                self._engines.clear()
            eng = EegnetEngine(self._dims, 1, B, device=self._arena.device)
            eng._owner_fwd_id = -1
            self._engines[B] = eng
        return eng

    # ------------------------------------------------------------ dropout masks
    def _draw_masks(self, B, device):
        """Keep-masks in the reference's draw order: mask1 (B,G,1,T/4) then mask2 (B,F2,1,T/32)
        from torch's global CPU generator (nn.Dropout on CPU draws exactly these)."""
        d = self._dims
        G, T4 = d.F1 * d.D, d.Samples // d.pool1
        T32 = T4 // d.pool2
        keep = 1.0 - d.dropoutRate
        if self._dropout2d:
            m1 = torch.empty(B, G, 1, 1).bernoulli_(keep).expand(B, G, 1, T4)
            m2 = torch.empty(B, d.F2, 1, 1).bernoulli_(keep).expand(B, d.F2, 1, T32)
        else:
            m1 = torch.empty(B, G, 1, T4).bernoulli_(keep)
            m2 = torch.empty(B, d.F2, 1, T32).bernoulli_(keep)
        to = lambda m, c, t: m.reshape(B, c, t).to(torch.uint8).contiguous().to(device, non_blocking=True)
        return to(m1, G, T4), to(m2, d.F2, T32)

    def _bump_bn_counters(self):
        for n in self._BN_NAMES:
            self.get_submodule(n).num_batches_tracked += 1

    # ------------------------------------------------------------ forward
    def _forward_cuda(self, x):
        if not torch.is_tensor(x) or not x.is_cuda:
            raise RuntimeError("eav_b200: forward() needs a CUDA tensor; there is no CPU fallback "
                               "(run the reference implementation for CPU execution)")
        d = self._dims
        if x.dim() == 4:
            if x.shape[1] != 1:
                raise ValueError(f"expected (B, 1, {d.Chans}, {d.Samples}), got {tuple(x.shape)}")
            x3 = x.reshape(x.shape[0], x.shape[2], x.shape[3])
        elif x.dim() == 3:
            x3 = x
        else:
            raise ValueError(f"expected (B, 1, Chans, Samples) or (B, Chans, Samples), got {tuple(x.shape)}")
        if x3.shape[1] != d.Chans or x3.shape[2] != d.Samples:
            raise ValueError(f"expected Chans={d.Chans}, Samples={d.Samples}, got {tuple(x3.shape)}")
        x3 = x3.to(torch.float32).contiguous()
        self._ensure_arena()
        B = x3.shape[0]
        masks = philox = None
        if self.training:
            self._bump_bn_counters()
            if d.dropoutRate > 0:
                if self.dropout_source == "torch_cpu":
                    masks = self._draw_masks(B, x3.device)
                else:
                    self._philox_step += 1
                    philox = (torch.initial_seed() & (2 ** 63 - 1), self._philox_step)
        plist = self._named_param_list()
        if torch.is_grad_enabled() and any(p.requires_grad for p in plist):
            return _EEGNetFn.apply(self, x3, masks, philox, *plist)
        eng = self._engine(B)
        self._fwd_id += 1
        eng._owner_fwd_id = self._fwd_id
        return eng.forward(x3, self._arena, self._bn_arena, bn_train=self.training,
                           mask1=masks[0] if masks else None, mask2=masks[1] if masks else None, philox=philox)


class ArenaAdam(torch.optim.Adam):
    """torch.optim.Adam (EEGNet_tor.py:82, CNN_EEG.py:86) whose per-parameter state IS the fused trainer's state:
    `exp_avg` / `exp_avg_sq` are views of the trainer's flat moment arenas and `step` mirrors its device-side
    step counter, so `trainer.optimizer.state_dict()` checkpoints the real Adam moments, `load_state_dict()`
    resumes them, and an edited `param_groups[0]['lr']` (LR schedulers) is what the next fused step uses."""

    def bind(self, core, layout, plist):
        self._core, self._layout, self._plist = core, layout, list(plist)
        self._alias()
        return self

    def _alias(self):
        c = self._core
        t = float(c.step_dev.item())
        for p, (name, off, shape) in zip(self._plist, self._layout):
            n = int(np.prod(shape))
            st = self.state[p]
            st["step"] = torch.tensor(t, dtype=torch.float32)
            st["exp_avg"] = c.exp_avg[0, off:off + n].view(shape)
            st["exp_avg_sq"] = c.exp_avg_sq[0, off:off + n].view(shape)

    def state_dict(self):
        if getattr(self, "_core", None) is not None:
            t = float(self._core.step_dev.item())
            for p in self._plist:
                self.state[p]["step"].fill_(t)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        c = getattr(self, "_core", None)
        if c is None:
            return
        steps = set()
        for p, (name, off, shape) in zip(self._plist, self._layout):
            st, n = self.state.get(p, {}), int(np.prod(shape))
            if "exp_avg" in st:
                c.exp_avg[0, off:off + n].copy_(st["exp_avg"].reshape(-1))
                c.exp_avg_sq[0, off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("eav_b200: the fused Adam keeps ONE step count for all parameters")
        if steps:
            c.step_dev.fill_(steps.pop())
        c.set_lr(float(self.param_groups[0]["lr"]))
        self._alias()

    def step(self, closure=None):
        """A manual optimizer.step() (p.grad populated by loss.backward() through the autograd Function) updates the
        same arenas the fused steps use; the shared step count advances with it."""
        out = super().step(closure)
        if getattr(self, "_core", None) is not None:
            self._core.step_dev += 1
        return out


def _as_rows(x, chans, samples):
    """numpy / tensor (N,C,T) or (N,1,C,T) -> float32 CPU tensor (N,C,T)."""
    t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
    t = t.detach().to("cpu", torch.float32)
    if t.dim() == 4 and t.shape[1] == 1:
        t = t[:, 0]
    if t.dim() != 3 or t.shape[1] != chans or t.shape[2] != samples:
        raise ValueError(f"expected (N, {chans}, {samples}) or (N, 1, {chans}, {samples}), got {tuple(t.shape)}")
    return t.contiguous()


class FusedTrainerMixin:
    """Resident-dataset, graph-captured training for ONE model behind the reference trainers."""

    def _setup_fused(self, model, tr_x, tr_y, te_x, te_y, lr=None, batch_size=None):
        d = model._dims
        self._model_ref = model
        trx, tex = _as_rows(tr_x, d.Chans, d.Samples), _as_rows(te_x, d.Chans, d.Samples)
        tr_y = torch.as_tensor(np.asarray(tr_y) if not torch.is_tensor(tr_y) else tr_y).long().cpu()
        te_y = torch.as_tensor(np.asarray(te_y) if not torch.is_tensor(te_y) else te_y).long().cpu()
        dev = next(model.parameters()).device
        self._n_train, self._n_test = trx.shape[0], tex.shape[0]
        x_all = torch.cat([trx, tex], 0).to(dev)
        y_all = torch.cat([tr_y, te_y], 0).to(dev)
        bs = batch_size if batch_size is not None else self.batch_size
        self._fused_bs = bs
        # index loaders: same sampler / RNG behaviour as the reference's loaders over (x, y)
        self._train_index_loader = DataLoader(TensorDataset(torch.arange(self._n_train)), batch_size=bs, shuffle=True)
        self._test_index_loader = DataLoader(TensorDataset(torch.arange(self._n_test)), batch_size=bs, shuffle=False)
        model._ensure_arena()
        self._core = SubjectBatchTrainer(d, 1, x_all, y_all, lr=lr if lr is not None else self.lr, max_batch=bs,
                                         params=model._arena, bn_state=model._bn_arena,
                                         seed=torch.initial_seed() & (2 ** 63 - 1))
        opt = getattr(self, "optimizer", None)
        if isinstance(opt, ArenaAdam):
            opt.bind(self._core, model._layout, model._named_param_list())

    def _index_batches(self, train=True):
        loader = self._train_index_loader if train else self._test_index_loader
        off = 0 if train else self._n_train
        dev = self._core.device
        for (rows,) in loader:
            yield (rows + off).to(torch.int32).to(dev, non_blocking=True)

    def _fused_train_step(self, rows):
        model = self._model_ref
        if not model._arena_ok():
            raise RuntimeError("eav_b200: model parameters were re-allocated after the trainer was built "
                               "(e.g. model.to(...)); rebuild the trainer")
        opt = getattr(self, "optimizer", None)
        if opt is not None and opt.param_groups and float(opt.param_groups[0]["lr"]) != self._core.lr:
            self._core.set_lr(float(opt.param_groups[0]["lr"]))       # LR scheduler / manual edit of param_groups
        masks = None
        if model.training:
            model._bump_bn_counters()
            if model._dims.dropoutRate > 0 and model.dropout_source == "torch_cpu":
                masks = model._draw_masks(rows.numel(), self._core.device)
        loss = self._core.train_step(rows, bn_train=model.training, masks=masks)
        if getattr(self, "record_losses", False):      # test / logging hook: keeps a device copy, no sync
            self.__dict__.setdefault("loss_history", []).append(loss[0].clone())
        return loss[0]

    def _fused_validate(self):
        tot_loss = torch.zeros((), dtype=torch.float64, device=self._core.device)
        tot_corr = torch.zeros((), dtype=torch.int64, device=self._core.device)
        nb = 0
        for rows in self._index_batches(train=False):
            loss, ncorrect, _ = self._core.eval_batch(rows)
            tot_loss += loss[0].double()
            tot_corr += ncorrect[0].long()
            nb += 1
        return float(tot_loss.item()), int(tot_corr.item()), nb
