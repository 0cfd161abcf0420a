"""Import the UNMODIFIED reference (nubcico/EAV at /root/reference) with the three
mechanical shims SURVEY.md §0 F1-F3 documents, so its CPU path can be executed in
this container to (a) pin the oracle restatement and (b) generate golden vectors.

TEST INFRASTRUCTURE ONLY.  Nothing here is shipped, nothing is copied from the
reference into the repository's history: only oracle/gen_golden.py, the `-m "not gpu"`
pin tests (skipped when the reference is absent) and `bench.py --impl reference` (the
CPU arm) import this module.

Shims (all mechanical, none changes arithmetic):
  F1  CNN_torch/EEGNet_tor.py:4 imports `Fusion.VIT_audio.Transformer_audio`,
      a package that is not in the repo -> pre-register empty stub modules.
  F2  CNN_torch/EEGNet_tor.py:92-93 uses TensorDataset/DataLoader without
      importing them -> inject both names into the module globals.
  F3  CNN_torch/EEGNet_tor.py:33-34,47-48 forward hooks return the renormed
      weight, which torch substitutes for the layer OUTPUT -> wrap every hook so
      it runs (the in-place renorm still happens) but returns None.
"""
import os
import sys
import types

# /root/reference in the build container; on the GPU box (where it does not exist) the verbatim copy of the
# path's four files that oracle/make_ref.py placed in the git-ignored oracle/_ref/.
_LOCAL_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("EAV_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/CNN_torch/EEGNet_tor.py") else _LOCAL_REF)


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "CNN_torch", "EEGNet_tor.py"))


def _ensure_path():
    for p in (REF_ROOT, os.path.join(REF_ROOT, "CNN_torch")):
        if p not in sys.path:
            sys.path.insert(0, p)


def load():
    """Returns a namespace with the reference modules (Dataload_eeg, EAV_datasplit,
    EEGNet_tor, CNN_EEG) imported from REF_ROOT."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _ensure_path()
    # F1: stub the missing package.
    for name in ("Fusion", "Fusion.VIT_audio", "Fusion.VIT_audio.Transformer_audio"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["Fusion.VIT_audio.Transformer_audio"].Trainer_uni = object
    import importlib
    import torch.utils.data as tud
    ns = types.SimpleNamespace()
    ns.Dataload_eeg = importlib.import_module("Dataload_eeg")
    ns.EAV_datasplit = importlib.import_module("EAV_datasplit")
    ns.EEGNet_tor = importlib.import_module("EEGNet_tor")
    ns.CNN_EEG = importlib.import_module("CNN_EEG")
    # F2: names the module forgot to import.
    ns.EEGNet_tor.TensorDataset = tud.TensorDataset
    ns.EEGNet_tor.DataLoader = tud.DataLoader
    return ns


def fix_hooks(model):
    """F3: make every forward hook of the reference EEGNet_tor return None."""
    for mod in (model.depthwiseConv, model.dense):
        for key, hook in list(mod._forward_hooks.items()):
            def wrapped(module, inputs, outputs, _h=hook):
                _h(module, inputs, outputs)
                return None
            mod._forward_hooks[key] = wrapped
    return model


def make_eegnet_tor(ns, *args, **kwargs):
    return fix_hooks(ns.EEGNet_tor.EEGNet_tor(*args, **kwargs))
