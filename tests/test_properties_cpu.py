"""Property tests (hypothesis) of the host-side integer logic against the CPU oracle and of the MAT v5 reader against
scipy.io.loadmat: ragged / degenerate inputs the fixed fixtures do not reach (bit-exact index work)."""
import numpy as np
import scipy.io
from hypothesis import given, settings, strategies as st

import eeg_oracle as O
from eav_b200.Dataload_eeg import epoch_slots
from eav_b200.EAV_datasplit import EAVDataSplit
from eav_b200 import mat_ingest as MI


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 9), min_size=1, max_size=60), st.integers(1, 5))
def test_epoch_slots_equal_oracle_plan(classes, n_sub):
    n = len(classes)
    label = np.zeros((10, n)); label[np.array(classes), np.arange(n)] = 1.0
    slot, y = epoch_slots(label, n_sub)
    keep, yo, src_trial, src_sub = O.epoch_plan(label, n_sub)
    assert np.array_equal(y, yo)
    assert np.array_equal(slot >= 0, keep)
    kept = np.nonzero(keep)[0]
    assert np.array_equal(slot[kept], np.arange(kept.size))              # kept trials keep their time order
    assert np.array_equal(src_trial, np.repeat(kept, n_sub)) and np.array_equal(src_sub, np.tile(np.arange(n_sub), kept.size))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 7), min_size=0, max_size=80), st.integers(0, 12))
def test_split_indices_equal_oracle(labels, h_idx):
    y = np.array(labels, dtype=np.int64)
    tr, te = EAVDataSplit(np.zeros((y.size, 1)), y).get_split_indices(h_idx=h_idx)
    tro, teo = O.split_indices(y, h_idx)
    assert np.array_equal(tr, tro) and np.array_equal(te, teo)
    used = np.concatenate([tr, te])
    assert np.array_equal(np.sort(used), np.nonzero(y < 5)[0])           # classes >= 5 are dropped, nothing else (F7)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 40), st.integers(1, 6), st.integers(1, 7), st.sampled_from(["f8", "f4", "i2", "u1"]),
       st.booleans(), st.integers(0, 2 ** 31 - 1))
def test_mat_reader_equals_loadmat(tmp_path_factory, t, c, k, dtype, compress, seed):
    rng = np.random.default_rng(seed)
    cnt = (rng.standard_normal((t, c, k)) * 50).astype(dtype)
    path = str(tmp_path_factory.mktemp("mat") / "a.mat")
    scipy.io.savemat(path, {"pad": np.arange(3.0), "seg": cnt}, do_compression=compress)
    arr, name, zero_copy = MI.read_mat_array(path, ("seg1", "seg"))
    ref = scipy.io.loadmat(path)["seg"]
    if ref.ndim == 2:                      # MATLAB drops trailing singleton dims beyond 2-D
        ref = ref.reshape(ref.shape + (1,) * (3 - ref.ndim))
    got = np.transpose(np.asarray(arr), tuple(reversed(range(arr.ndim))))
    assert name == "seg" and got.dtype == ref.dtype and np.array_equal(got.reshape(ref.shape), ref)
