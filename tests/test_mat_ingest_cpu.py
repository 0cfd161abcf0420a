"""CPU tests of eav_b200.mat_ingest: the MAT-file v5 reader against scipy.io.loadmat (bit-exact, the reference's
loader at Dataload_eeg.py:70-79) and the prefetching iterator in host-only mode."""
import numpy as np
import pytest
import scipy.io

from eav_b200 import mat_ingest as MI


@pytest.mark.parametrize("compress", [False, True])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int16])
@pytest.mark.parametrize("name", ["seg", "seg1"])
def test_reader_matches_loadmat(tmp_path, compress, dtype, name):
    rng = np.random.default_rng(3)
    cnt = (rng.standard_normal((500, 6, 7)) * 100).astype(dtype)            # (Time, Channels, Trials)
    path = str(tmp_path / "x.mat")
    scipy.io.savemat(path, {"other": np.arange(5.0), name: cnt, "zz": "text"}, do_compression=compress)
    arr, got_name, zero_copy = MI.read_mat_array(path, (name,))
    ref = scipy.io.loadmat(path)[name]
    assert got_name == name and zero_copy == (not compress)
    assert arr.dtype == ref.dtype and arr.shape == (7, 6, 500)
    assert np.array_equal(np.transpose(np.asarray(arr), (2, 1, 0)), ref)    # arr[k, c, t] == ref[t, c, k]


def test_reader_rejects_what_it_does_not_parse(tmp_path):
    path = str(tmp_path / "c.mat")
    scipy.io.savemat(path, {"seg": np.ones((3, 3)) + 1j})
    with pytest.raises(KeyError):
        MI.read_mat_array(path, ("seg",))                                   # complex -> not returned
    with pytest.raises(KeyError):
        MI.read_mat_array(path, ("nope",))
    bad = tmp_path / "h.mat"
    bad.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 200)
    with pytest.raises(MI.MatFormatError):
        MI.read_mat_array(str(bad), ("seg",))


def _write_subject(root, s, rng, compress=False, name="seg"):
    folder = root / f"subject{s:02d}" / "EEG"
    folder.mkdir(parents=True)
    cnt = rng.standard_normal((200, 5, 12))                                 # (Time, Channels, Trials) float64
    label = np.zeros((10, 12)); label[rng.integers(0, 10, 12), np.arange(12)] = 1
    scipy.io.savemat(str(folder / f"subject{s:02d}_eeg.mat"), {name: cnt}, do_compression=compress)
    scipy.io.savemat(str(folder / f"subject{s:02d}_eeg_label.mat"), {"label": label})
    return cnt, label


def test_load_subject_and_seg1_preference(tmp_path):
    rng = np.random.default_rng(5)
    cnt, label = _write_subject(tmp_path, 3, rng, name="seg1")
    raw, lab, zero_copy = MI.load_subject_mat(str(tmp_path), 3)
    assert zero_copy and np.array_equal(np.transpose(np.asarray(raw), (2, 1, 0)), cnt) and np.array_equal(lab, label)
    with pytest.raises(FileNotFoundError):
        MI.load_subject_mat(str(tmp_path), 4)


def test_prefetcher_host_mode_order_and_values(tmp_path):
    rng = np.random.default_rng(9)
    truth = {s: _write_subject(tmp_path, s, rng, compress=(s % 2 == 0)) for s in (1, 2, 3, 4, 5)}
    seen = []
    for s, raw, label in MI.SubjectPrefetcher(str(tmp_path), [1, 2, 3, 4, 5], device=None, depth=2, chunk_trials=5):
        cnt, lab = truth[s]
        assert raw.dtype.is_floating_point and tuple(raw.shape) == (12, 5, 200)
        assert np.array_equal(raw.numpy(), np.transpose(cnt, (2, 1, 0)).astype(np.float32))
        assert np.array_equal(label, lab)
        seen.append(s)
    assert seen == [1, 2, 3, 4, 5]


def test_prefetcher_propagates_errors(tmp_path):
    rng = np.random.default_rng(1)
    _write_subject(tmp_path, 1, rng)
    with pytest.raises(FileNotFoundError):
        list(MI.SubjectPrefetcher(str(tmp_path), [1, 2], device=None))


def _mat5_double_stored_as_int16(path, name, values):
    """A MAT-file v5 whose mxDOUBLE array `name` has its real part STORED as miINT16 (MATLAB's numeric data
    compression): hand-assembled, since scipy.io.savemat never writes that form."""
    import struct
    vals = np.asarray(values)
    dims = vals.shape
    body = struct.pack("<II", 6, 8) + struct.pack("<II", 6, 0)                     # array flags: class mxDOUBLE (6)
    body += struct.pack("<II", 5, 4 * len(dims)) + struct.pack("<%di" % len(dims), *dims)
    body += b"\0" * ((-4 * len(dims)) % 8)
    nm = name.encode()
    body += struct.pack("<II", 1, len(nm)) + nm + b"\0" * ((-len(nm)) % 8)
    data = np.asfortranarray(vals).astype("<i2").tobytes(order="F")
    body += struct.pack("<II", 3, len(data)) + data + b"\0" * ((-len(data)) % 8)   # miINT16 payload
    head = b"MATLAB 5.0 MAT-file, hand-made".ljust(116) + b"\0" * 8 + struct.pack("<H", 0x0100) + b"IM"
    with open(path, "wb") as f:
        f.write(head + struct.pack("<II", 14, len(body)) + body)


def test_double_array_stored_as_int16_matches_loadmat_defaults(tmp_path):
    """ADVICE r1 asked for float64 here; scipy.io.loadmat with its defaults (what the reference calls,
    Dataload_eeg.py:70) returns the STORAGE dtype for such an array, and the reader must agree with loadmat."""
    vals = np.arange(4 * 3 * 2).reshape(4, 3, 2) - 7
    folder = tmp_path / "subject01" / "EEG"
    folder.mkdir(parents=True)
    path = str(folder / "subject01_eeg.mat")
    _mat5_double_stored_as_int16(path, "seg", vals)
    ref = scipy.io.loadmat(path)["seg"]
    assert np.array_equal(ref, vals)
    arr, _, zero_copy = MI.read_mat_array(path, ("seg",))
    assert zero_copy and arr.dtype == ref.dtype                                # int16, like loadmat's default
    assert np.array_equal(np.transpose(np.asarray(arr), (2, 1, 0)), ref)
    assert scipy.io.loadmat(path, mat_dtype=True)["seg"].dtype == np.float64   # (the class dtype needs mat_dtype=True)
