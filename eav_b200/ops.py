"""Thin host-side wrappers over the C ABI: torch owns device memory and streams, the
kernels in libeav_b200.so do all the arithmetic.  No CPU path exists here on purpose.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import (EAV_DROPOUT_MASK, EAV_DROPOUT_NONE, EAV_DROPOUT_PHILOX, EAV_DROPOUT_PHILOX_2D, EAV_VARIANT_CNN,
                   EAV_VARIANT_TOR,
                   EegnetCfg, PreprocCfg)

TOR_PARAM_NAMES = ("firstConv.weight", "firstBN.weight", "firstBN.bias", "depthwiseConv.weight",
                   "depthwiseBN.weight", "depthwiseBN.bias", "separableConv.weight", "separableBN.weight",
                   "separableBN.bias", "dense.weight", "dense.bias")
CNN_PARAM_NAMES = ("block1.0.weight", "block1.1.weight", "block1.1.bias", "block1.2.weight", "block1.3.weight",
                   "block1.3.bias", "block2.0.weight", "block2.1.weight", "block2.2.weight", "block2.2.bias",
                   "classifier.weight", "classifier.bias")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    """The CURRENT device's current stream.  The C ABI launches on the current device (kernel attributes, SM count
    and the launch itself are per device), so every caller first makes the device that owns its tensors current
    (`with _on(device):`) -- a Trainer_uni(device='cuda:1') in a process whose current device is 0 then launches
    on GPU 1's stream with GPU 1's pointers."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _norm_device(device=None):
    """torch.device with an explicit index ('cuda' -> the current device), so tensor.device comparisons are exact."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _on(device):
    """Context manager: `device` is the current CUDA device inside the block."""
    return torch.cuda.device(device)


def _chk_cuda(t, dtype, name, device=None):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (eav_b200 has no CPU path)")
    if device is not None and t.device != device:
        raise RuntimeError(f"{name} lives on {t.device}, the engine on {device}")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


@dataclass
class EegnetDims:
    """Model hyper-parameters in the reference's vocabulary (EEGNet_tor.py:16-17 / CNN_EEG.py:12-13)."""
    nb_classes: int
    Chans: int = 30
    Samples: int = 500
    dropoutRate: float = 0.5
    kernLength: int = 300
    F1: int = 8
    D: int = 8
    F2: int = 64
    norm_rate: float = 1.0
    variant: int = EAV_VARIANT_TOR
    kernLength2: int = 16
    pool1: int = 4
    pool2: int = 8
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1
    dropout2d: bool = False     # dropoutType != 'Dropout' (EEGNet_tor.py:21): nn.Dropout2d, whole channels dropped

    @property
    def philox_mode(self):
        """The on-device dropout mode of this model (element-wise or per (sample, channel) row)."""
        return EAV_DROPOUT_PHILOX_2D if self.dropout2d else EAV_DROPOUT_PHILOX

    @property
    def n_bn(self):
        return 2 * (self.F1 + self.F1 * self.D + self.F2)

    def cfg(self, n_models=1, batch=1, bn_train=False, dropout_mode=EAV_DROPOUT_NONE, param_stride=None,
            bn_stride=None, seed=0, step=0, step_ptr=0, dp_world=1) -> EegnetCfg:
        c = EegnetCfg()
        c.n_models, c.batch = n_models, batch
        c.chans, c.samples, c.kern_len = self.Chans, self.Samples, self.kernLength
        c.F1, c.D, c.F2, c.kern_len2 = self.F1, self.D, self.F2, self.kernLength2
        c.pool1, c.pool2, c.n_classes = self.pool1, self.pool2, self.nb_classes
        c.variant, c.bn_train, c.dropout_mode = self.variant, int(bool(bn_train)), dropout_mode
        c.dropout_p, c.bn_eps, c.bn_momentum = self.dropoutRate, self.bn_eps, self.bn_momentum
        c.norm_rate = self.norm_rate if self.variant == EAV_VARIANT_TOR else 0.0
        c.seed, c.step, c.step_device_ptr = seed, step, step_ptr
        c.dp_world = dp_world
        c.param_stride = param_stride if param_stride is not None else 2 ** 31 - 1
        c.bn_stride = bn_stride if bn_stride is not None else 2 ** 31 - 1
        return c

    def param_layout(self):
        """(n_params, [(name, offset, shape)]) in the reference's parameter order."""
        offs = (ctypes.c_int64 * 12)()
        c = self.cfg()
        n = _lib.load().eav_eegnet_param_layout(ctypes.byref(c), offs)
        if n < 0:
            raise RuntimeError(f"eav_eegnet_param_layout failed: {_lib.last_error()}")
        G, T32 = self.F1 * self.D, self.Samples // self.pool1 // self.pool2
        feat = self.F2 * T32
        if self.variant == EAV_VARIANT_TOR:
            shapes = [(self.F1, 1, 1, self.kernLength), (self.F1,), (self.F1,), (G, 1, self.Chans, 1), (G,), (G,),
                      (self.F2, G, 1, self.kernLength2), (self.F2,), (self.F2,), (self.nb_classes, feat),
                      (self.nb_classes,)]
            names = TOR_PARAM_NAMES
        else:
            shapes = [(self.F1, 1, 1, self.kernLength), (self.F1,), (self.F1,), (G, 1, self.Chans, 1), (G,), (G,),
                      (G, 1, 1, self.kernLength2), (self.F2, G, 1, 1), (self.F2,), (self.F2,),
                      (self.nb_classes, feat), (self.nb_classes,)]
            names = CNN_PARAM_NAMES
        return int(n), [(nm, int(offs[i]), sh) for i, (nm, sh) in enumerate(zip(names, shapes))]

    def bn_layout(self):
        """[(buffer suffix, offset, size)] of one model's slice of bn_state: rm1 rv1 rm2 rv2 rm3 rv3."""
        G = self.F1 * self.D
        out, o = [], 0
        for i, ch in enumerate((self.F1, G, self.F2)):
            out.append((i, "running_mean", o, ch)); o += ch
            out.append((i, "running_var", o, ch)); o += ch
        return out


class EegnetEngine:
    """Owns the workspace for a fixed (dims, n_models, batch) and issues the C-ABI calls."""

    def __init__(self, dims: EegnetDims, n_models: int, batch: int, device=None):
        _lib.require_device()
        self.lib = _lib.load()
        self.dims, self.M, self.B = dims, n_models, batch
        self.device = _norm_device(device)
        self.n_params, self.layout = dims.param_layout()
        c = dims.cfg(n_models, batch, param_stride=self.n_params, bn_stride=dims.n_bn)
        nbytes = self.lib.eav_eegnet_workspace_bytes(ctypes.byref(c))
        if nbytes == 0:
            raise RuntimeError(f"eav_eegnet_workspace_bytes failed: {_lib.last_error()}")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.ws_bytes = nbytes

    def saved(self, name):
        """Inspection: a float32 view of one saved activation / scratch tensor in the workspace."""
        d, n = self.dims, self.M * self.B
        G, T4 = d.F1 * d.D, d.Samples // d.pool1
        shapes = {"y1": (n, d.F1, d.Chans, d.Samples), "y2": (n, G, d.Samples), "d1": (n, G, T4),
                  "y3d": (n, G, T4), "y3": (n, d.F2, T4), "feat": (n, d.F2 * (T4 // d.pool2)),
                  "probs": (n, d.nb_classes), "dz": (n, d.nb_classes), "dz3": (n, d.F2, T4), "dd1": (n, G, T4),
                  "dy3d": (n, G, T4), "dz2": (n, G, d.Samples), "dz1": (n, d.F1, d.Chans, d.Samples),
                  "bnf1": (self.M, d.F1, 4), "bnf2": (self.M, G, 4), "bnf3": (self.M, d.F2, 4)}
        offs = (ctypes.c_size_t * 16)()
        c = self.dims.cfg(self.M, self.B, param_stride=self.n_params, bn_stride=self.dims.n_bn)
        _lib.check(self.lib.eav_eegnet_workspace_offsets(ctypes.byref(c), offs), "eav_eegnet_workspace_offsets")
        idx = list(shapes).index(name)
        numel = int(np.prod(shapes[name]))
        return self.workspace[offs[idx]:offs[idx] + 4 * numel].view(torch.float32).view(shapes[name])

    def _cfg(self, params, bn_state, bn_train, dropout_mode, seed=0, step=0):
        pstride = params.stride(0) if params.dim() == 2 else self.n_params
        bstride = bn_state.stride(0) if (bn_state is not None and bn_state.dim() == 2) else self.dims.n_bn
        return self.dims.cfg(self.M, self.B, bn_train, dropout_mode, pstride, bstride, seed, step)

    def forward(self, x, params, bn_state, bn_train=False, x_index=None, mask1=None, mask2=None, philox=None,
                out=None):
        """x [rows][C][T] f32; params [M][P]; bn_state [M][n_bn]; returns out [M*B][nb_classes]."""
        dv = self.device
        _chk_cuda(x, torch.float32, "x", dv); _chk_cuda(params, torch.float32, "params", dv)
        _chk_cuda(bn_state, torch.float32, "bn_state", dv); _chk_cuda(x_index, torch.int32, "x_index", dv)
        _chk_cuda(mask1, torch.uint8, "mask1", dv); _chk_cuda(mask2, torch.uint8, "mask2", dv)
        mode = EAV_DROPOUT_NONE
        seed = step = 0
        if bn_train and self.dims.dropoutRate > 0:
            if mask1 is not None:
                mode = EAV_DROPOUT_MASK
            elif philox is not None:
                mode, (seed, step) = self.dims.philox_mode, philox
            else:
                raise ValueError("train-mode forward needs dropout masks or a philox (seed, step)")
        n = self.M * self.B
        if x_index is None and x.shape[0] < n:
            raise ValueError(f"x has {x.shape[0]} rows, need {n}")
        if out is None:
            out = torch.empty(n, self.dims.nb_classes, dtype=torch.float32, device=self.device)
        c = self._cfg(params, bn_state, bn_train, mode, seed, step)
        self._last = (c, mode)
        with _on(self.device):
            _lib.check(self.lib.eav_eegnet_forward(ctypes.byref(c), _ptr(x), _ptr(x_index), _ptr(params), _ptr(bn_state),
                                                   _ptr(mask1), _ptr(mask2), _ptr(out), _ptr(self.workspace),
                                                   self.ws_bytes, _stream()), "eav_eegnet_forward")
        return out

    def loss(self, out, targets, x_index=None, want_grad=True):
        _chk_cuda(out, torch.float32, "out", self.device); _chk_cuda(targets, torch.int64, "targets", self.device)
        loss = torch.empty(self.M, dtype=torch.float32, device=self.device)
        ncorrect = torch.empty(self.M, dtype=torch.int32, device=self.device)
        dout = torch.empty_like(out) if want_grad else None
        c = self.dims.cfg(self.M, self.B)
        with _on(self.device):
            _lib.check(self.lib.eav_eegnet_loss(ctypes.byref(c), _ptr(out), _ptr(targets), _ptr(x_index), _ptr(loss),
                                                _ptr(dout), _ptr(ncorrect), _stream()), "eav_eegnet_loss")
        return loss, dout, ncorrect

    def backward(self, x, params, dout, grads=None, x_index=None, mask1=None, mask2=None):
        """Gradient of the LAST forward() (same cfg / workspace). Returns grads [M][P]."""
        c, _ = self._last
        _chk_cuda(dout, torch.float32, "dout", self.device)
        if grads is None:
            grads = torch.zeros(self.M, params.stride(0) if params.dim() == 2 else self.n_params,
                                dtype=torch.float32, device=self.device)
        with _on(self.device):
            _lib.check(self.lib.eav_eegnet_backward(ctypes.byref(c), _ptr(x), _ptr(x_index), _ptr(params), _ptr(dout),
                                                    _ptr(mask1), _ptr(mask2), _ptr(grads), _ptr(self.workspace),
                                                    self.ws_bytes, _stream()), "eav_eegnet_backward")
        return grads


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8):
    for t, nm in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _chk_cuda(t, torch.float32, nm)
    with _on(params.device):
        _lib.check(_lib.load().eav_adam_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(),
                                             int(step), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                             _stream()), "eav_adam_step")


def renorm_rows(w2d, maxnorm):
    _chk_cuda(w2d, torch.float32, "w")
    rows, ln = w2d.shape
    with _on(w2d.device):
        _lib.check(_lib.load().eav_renorm_rows(_ptr(w2d), rows, ln, w2d.stride(0), float(maxnorm), _stream()),
                   "eav_renorm_rows")


def measure_fp32_peak(mode=0) -> float:
    """Measured fp32 TFLOP/s of this device.  mode 0: register-resident FFMA loop with immediate /
    uniform operands (the pipe peak); 1: 8x8 register outer product with scalar FFMA; 2: the same with
    one operand in a uniform register; 3: the same with packed FFMA2 (fma.rn.f32x2)."""
    _lib.require_device()
    v = ctypes.c_double(0.0)
    _lib.check(_lib.load().eav_measure_fp32_peak_mode(int(mode), ctypes.byref(v), _stream()), "eav_measure_fp32_peak_mode")
    return v.value


def tc_probe(image, M, N, ksteps, reps, a, b, n_acc=1, a_bits=0, b_bits=0):
    """Runs reps x ksteps tcgen05.mma.kind::tf32 on a shared-memory image (1-D fp32 CUDA tensor) with the
    operand descriptors a, b = (byte offset, LBO, SBO, major, byte step per k).  Returns (D[128, N], cycles)."""
    _lib.require_device()
    _chk_cuda(image, torch.float32, "image")
    d = torch.zeros(128, N, dtype=torch.float32, device=image.device)
    cyc = torch.zeros(1, dtype=torch.int64, device=image.device)
    _lib.check(_lib.load().eav_tc_probe(_ptr(image), image.numel(), M, N, ksteps, reps, n_acc, *[int(v) for v in a],
                                        *[int(v) for v in b], int(a_bits), int(b_bits), _ptr(d), _ptr(cyc), _stream()), "eav_tc_probe")
    torch.cuda.synchronize()
    return d, int(cyc.item())


# ---------------------------------------------------------------------------------------
# preprocessing
# ---------------------------------------------------------------------------------------
class PreprocEngine:
    """Decimate + band-pass + epoch a batch of subjects on the GPU (Dataload_eeg.py:85-152)."""

    def __init__(self, n_subjects, n_trials=200, n_chans=30, trial_len=10000, down=5, n_taps=101, n_sections=5,
                 n_sub=4, raw_dtype=torch.float32, device=None, order=0):
        """order 0: decimate then band-pass at fs_target (Dataload_eeg.py); 1: band-pass at fs_orig then decimate
        (legacy CNN_EEG_tf.py:64-75,182-189; `sos` must then be designed for fs_orig)."""
        _lib.require_device()
        self.lib = _lib.load()
        self.device = _norm_device(device)
        c = PreprocCfg()
        c.n_subjects, c.n_trials, c.n_chans, c.trial_len = n_subjects, n_trials, n_chans, trial_len
        c.down, c.n_taps, c.n_sections, c.n_sub = down, n_taps, n_sections, n_sub
        c.raw_is_f64 = int(raw_dtype == torch.float64)
        c.order = int(order)
        self.cfg = c
        self.raw_dtype = raw_dtype
        nbytes = self.lib.eav_preproc_workspace_bytes(ctypes.byref(c))
        if nbytes == 0:
            raise RuntimeError(f"eav_preproc_workspace_bytes failed: {_lib.last_error()}")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.ws_bytes = nbytes
        self.ep_len = trial_len // down // n_sub
        self.n_dec = n_trials * (trial_len // down)

    def run(self, raw, taps, sos, epoch_slot, n_epochs_out, epochs=None, want_dec=False):
        """raw [S][trials][ch][time]; taps f64[n_taps], sos f64[n_sections][6] (host numpy);
        epoch_slot i32 [S][trials] (device).  Returns epochs [S][n_epochs_out][ch][ep_len] f32
        (and dec [S][ch][n_dec] f32 when want_dec)."""
        c = self.cfg
        _chk_cuda(raw, self.raw_dtype, "raw", self.device); _chk_cuda(epoch_slot, torch.int32, "epoch_slot", self.device)
        if tuple(raw.shape) != (c.n_subjects, c.n_trials, c.n_chans, c.trial_len):
            raise ValueError(f"raw has shape {tuple(raw.shape)}")
        taps = np.ascontiguousarray(taps, dtype=np.float64)
        sos = np.ascontiguousarray(sos, dtype=np.float64)
        if taps.size != c.n_taps or sos.shape != (c.n_sections, 6):
            raise ValueError("taps/sos do not match the engine configuration")
        if epochs is None:
            epochs = torch.empty(c.n_subjects, n_epochs_out, c.n_chans, self.ep_len, dtype=torch.float32,
                                 device=self.device)
        dec = torch.empty(c.n_subjects, c.n_chans, self.n_dec, dtype=torch.float32, device=self.device) if want_dec else None
        dp = ctypes.POINTER(ctypes.c_double)
        with _on(self.device):
            _lib.check(self.lib.eav_preproc_run(ctypes.byref(c), _ptr(raw), taps.ctypes.data_as(dp), sos.ctypes.data_as(dp),
                                                _ptr(epoch_slot), int(n_epochs_out), _ptr(epochs), _ptr(dec),
                                                _ptr(self.workspace), self.ws_bytes, _stream()), "eav_preproc_run")
        return (epochs, dec) if want_dec else epochs
