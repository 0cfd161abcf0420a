"""CPU checks of the host-side mirrors of the reference interface: EAVDataSplit index logic
(bit-exact vs reference golden), module construction (same submodules / keys / default
initialisation stream as the reference), epoch-slot planning, loud failure without a GPU."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("tag", ["shipped", "remap"])
def test_eavdatasplit_bit_exact(golden, tag):
    from eav_b200.EAV_datasplit import EAVDataSplit
    g = golden(f"split_{tag}.npz")
    y = g["y"]
    x = np.arange(y.size, dtype=np.float64).reshape(-1, 1, 1) * np.ones((1, 2, 3))
    for h in (40, 56):
        sp = EAVDataSplit(x, y)
        trx, try_, tex, tey = sp.get_split(h_idx=h)
        assert np.array_equal(trx[:, 0, 0].astype(np.int64), g[f"tr_idx_{h}"])
        assert np.array_equal(tex[:, 0, 0].astype(np.int64), g[f"te_idx_{h}"])
        assert np.array_equal(try_, g[f"tr_y_{h}"]) and np.array_equal(tey, g[f"te_y_{h}"])
        tri, tei = sp.get_split_indices(h)
        assert np.array_equal(tri, g[f"tr_idx_{h}"]) and np.array_equal(tei, g[f"te_idx_{h}"])
    # np.squeeze drops ALL singleton dims like the reference
    sp = EAVDataSplit(np.zeros((y.size, 1, 30, 5)), y)
    assert sp.get_split(56)[0].shape[1:] == (30, 5)
    tl, te = EAVDataSplit(np.zeros((y.size, 3)), y, batch_size=7).get_loaders()
    assert tl.batch_size == 7 and len(te.dataset) == len(sp.get_split()[2])


def test_epoch_slots_match_reference_plan(golden):
    from eav_b200.Dataload_eeg import epoch_slots
    g = golden("segment_plan.npz")
    slot, y = epoch_slots(g["label"].astype(np.float64))
    assert np.array_equal(y, g["y"])
    kept = np.nonzero(slot >= 0)[0]
    assert np.array_equal(slot[kept], np.arange(kept.size))
    assert np.array_equal(np.repeat(kept, 4), g["src_trial"])


def test_eegnet_tor_construction_matches_reference_init(golden):
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    g = golden("eegnet_tor_b8.npz")
    torch.manual_seed(3)
    m = EEGNet_tor(5)
    sd = m.state_dict()
    ref_keys = [k[6:] for k in g.files if k.startswith("init::")]
    assert list(sd.keys()) == ref_keys
    for k in ("firstConv.weight", "depthwiseConv.weight", "separableConv.weight", "dense.weight", "dense.bias"):
        assert np.array_equal(sd[k].numpy(), g["init::" + k]), k
    assert sum(p.numel() for p in m.parameters()) == 74933


def test_cnn_eeg_construction_matches_reference_init(golden):
    from eav_b200.CNN_torch.CNN_EEG import EEGNet
    g = golden("cnn_eeg_default.npz")
    torch.manual_seed(4)
    m = EEGNet(nb_classes=4, Chans=64, Samples=128, dropoutRate=0.25)
    sd = m.state_dict()
    assert list(sd.keys()) == [k[6:] for k in g.files if k.startswith("init::")]
    for k in ("block1.0.weight", "block1.2.weight", "block2.0.weight", "block2.1.weight", "classifier.weight", "classifier.bias"):
        assert np.array_equal(sd[k].numpy(), g["init::" + k]), k
    # side effects of the reference's construction-time dry run (CNN_EEG.py:48-53)
    assert int(sd["block1.1.num_batches_tracked"]) == 1 and float(sd["block1.1.running_var"][0]) == pytest.approx(0.9)


def test_forward_on_cpu_fails_loudly():
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    m = EEGNet_tor(5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 1, 30, 500))


def test_dataload_signature_and_aliases():
    import inspect
    from eav_b200.Dataload_eeg import DataLoadEEG
    sig = inspect.signature(DataLoadEEG.__init__)
    assert list(sig.parameters)[1:] == ["subject", "band", "fs_orig", "fs_target", "parent_directory"]
    d = DataLoadEEG()
    assert d.band == [0.3, 50] and d.fs_orig == 500 and d.fs_target == 100
    for name in ("load_mat_data", "downsampling", "bandpass_filter", "segment_and_select_classes", "prepare_data",
                 "data_mat", "bandpass", "data_div", "data_prepare"):
        assert callable(getattr(d, name))
    assert d.prepare_data() == (None, None)      # missing files: prints the error, returns the empty placeholders
