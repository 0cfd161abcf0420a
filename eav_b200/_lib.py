"""ctypes binding of libeav_b200.so (include/eav_b200.h).

There is NO CPU fallback: if the shared library is missing the import of any compute
entry point raises, and every call checks its return code and raises RuntimeError with
eav_last_error_string().  Build the library with `python -m eav_b200.build`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint32, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeav_b200.so")

EAV_VARIANT_TOR, EAV_VARIANT_CNN = 0, 1
EAV_DROPOUT_NONE, EAV_DROPOUT_MASK, EAV_DROPOUT_PHILOX, EAV_DROPOUT_PHILOX_2D = 0, 1, 2, 3


class PreprocCfg(Structure):
    _fields_ = [(n, c_int32) for n in ("n_subjects", "n_trials", "n_chans", "trial_len", "down", "n_taps",
                                       "n_sections", "n_sub", "raw_is_f64", "order")]


class EegnetCfg(Structure):
    _fields_ = ([(n, c_int32) for n in ("n_models", "batch", "chans", "samples", "kern_len", "F1", "D", "F2",
                                        "kern_len2", "pool1", "pool2", "n_classes", "variant", "bn_train",
                                        "dropout_mode", "param_stride", "bn_stride")]
                + [(n, c_float) for n in ("dropout_p", "bn_eps", "bn_momentum", "norm_rate")]
                + [("seed", c_uint64), ("step", c_uint64), ("step_device_ptr", c_uint64)]
                + [("dp_world", c_int32), ("reserved", c_int32)])


class ShallowCfg(Structure):
    _fields_ = ([(n, c_int32) for n in ("batch", "chans", "samples", "n_filters", "kern", "n_layers", "ffn", "pool",
                                        "stride", "n_classes", "bn_train", "dropout_mode")]
                + [(n, c_float) for n in ("dropout_p", "bn_eps", "bn_momentum", "ln_eps")])


# every symbol include/eav_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "eav_last_error_string": (c_char_p, []),
    "eav_abi_version": (c_int, []),
    "eav_launch_count": (c_uint64, []),
    "eav_check_device": (c_int, []),
    "eav_preproc_workspace_bytes": (c_size_t, [POINTER(PreprocCfg)]),
    "eav_preproc_run": (c_int, [POINTER(PreprocCfg), c_void_p, POINTER(c_double), POINTER(c_double), c_void_p,
                                c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "eav_eegnet_param_layout": (c_int64, [POINTER(EegnetCfg), POINTER(c_int64)]),
    "eav_eegnet_workspace_bytes": (c_size_t, [POINTER(EegnetCfg)]),
    "eav_eegnet_workspace_offsets": (c_int, [POINTER(EegnetCfg), POINTER(c_size_t)]),
    "eav_eegnet_forward": (c_int, [POINTER(EegnetCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "eav_eegnet_apply_hooks": (c_int, [POINTER(EegnetCfg), c_void_p, c_void_p]),
    "eav_eegnet_loss": (c_int, [POINTER(EegnetCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "eav_eegnet_backward": (c_int, [POINTER(EegnetCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_size_t, c_void_p]),
    "eav_eegnet_stage_count": (c_int, []),
    "eav_eegnet_stage_forward_end": (c_int, []),
    "eav_eegnet_stage_name": (c_char_p, [c_int]),
    "eav_eegnet_run_stage": (c_int, [POINTER(EegnetCfg), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "eav_eegnet_stage_allreduce": (c_int, [POINTER(EegnetCfg), c_int, POINTER(c_size_t), POINTER(c_size_t)]),
    "eav_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_float,
                              c_float, c_void_p]),
    "eav_adam_step_graph": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float,
                                    c_float, c_float, c_void_p]),
    "eav_epoch_schedule": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_int64, c_uint64, c_void_p,
                                   c_void_p]),
    "eav_epoch_accumulate": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "eav_epoch_commit": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p,
                                 c_void_p]),
    "eav_shallow_param_layout": (c_int64, [POINTER(ShallowCfg), POINTER(c_int64), POINTER(c_int64)]),
    "eav_shallow_workspace_bytes": (c_size_t, [POINTER(ShallowCfg)]),
    "eav_shallow_forward": (c_int, [POINTER(ShallowCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "eav_shallow_backward": (c_int, [POINTER(ShallowCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "eav_epoch_commit_val": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p]),
    "eav_renorm_rows": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p]),
    "eav_measure_fp32_peak": (c_int, [POINTER(c_double), c_void_p]),
    "eav_measure_fp32_peak_outer": (c_int, [POINTER(c_double), c_void_p]),
    "eav_measure_fp32_peak_mode": (c_int, [c_int, POINTER(c_double), c_void_p]),
    "eav_peer_exchange_bytes": (c_size_t, [c_size_t]),
    "eav_peer_allreduce": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_size_t, c_uint32, c_void_p,
                                  c_uint32, c_void_p]),
    "eav_tc_probe": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int] + [c_int] * 12 + [c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen libeav_b200.so and bind every declared symbol; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python -m eav_b200.build`). "
            "eav_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().eav_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")


def require_device():
    """Fail loudly unless a CUDA sm_100 device is current."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("eav_b200 needs a CUDA (B200, sm_100a) device; there is no CPU fallback")
    check(load().eav_check_device(), "eav_check_device")
