"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/eav_b200.h declares; argument errors are reported, not crashed on.
No compute entry point is called (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from eav_b200.build import build
    build()
    from eav_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from eav_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "eav_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(eav_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_layout(lib):
    from eav_b200.ops import EegnetDims
    from eav_b200._lib import EAV_VARIANT_CNN
    assert lib.eav_abi_version() == 1
    n, layout = EegnetDims(5).param_layout()
    assert n == 74933                                   # 74 933 parameters
    assert [nm for nm, _, _ in layout][0] == "firstConv.weight" and layout[-1][0] == "dense.bias"
    n2, _ = EegnetDims(4, Chans=64, Samples=128, kernLength=64, D=2, F2=16, variant=EAV_VARIANT_CNN).param_layout()
    assert n2 == 2388                                   # CNN_EEG defaults
    n3, _ = EegnetDims(5, Chans=30, Samples=500, kernLength=300, D=8, F2=64, variant=EAV_VARIANT_CNN).param_layout()
    assert n3 == 14517


def test_argument_errors_are_reported(lib):
    from eav_b200._lib import EegnetCfg, PreprocCfg, last_error
    c = EegnetCfg()
    assert lib.eav_eegnet_workspace_bytes(ctypes.byref(c)) == 0
    assert "positive" in last_error()
    p = PreprocCfg()
    assert lib.eav_preproc_workspace_bytes(ctypes.byref(p)) == 0
    from eav_b200.ops import EegnetDims
    cfg = EegnetDims(5).cfg(42, 32, param_stride=74933, bn_stride=272)
    assert lib.eav_eegnet_workspace_bytes(ctypes.byref(cfg)) > 10 ** 9   # ~2 GB of saved activations at 1344 samples
    cfg.param_stride = 10
    assert lib.eav_eegnet_workspace_bytes(ctypes.byref(cfg)) == 0 and "param_stride" in last_error()


def test_no_cpu_fallback():
    import torch
    from eav_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_device()
