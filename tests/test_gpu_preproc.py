"""-m gpu parity tests of the preprocessing CUDA path (FIR decimate -> SOS scan -> epoch
scatter, through the C ABI) against the reference's golden vectors and the CPU oracle.
Gate: epochs within 1e-5 of the per-channel RMS of the reference's float64
result; epoch / label indexing bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _slots(keep):
    """epoch_slot for one subject: running index over kept trials, -1 for dropped ones."""
    slot = np.full(keep.shape, -1, dtype=np.int32)
    slot[keep] = np.arange(int(keep.sum()), dtype=np.int32)
    return slot


def _chan_rel_err(got, ref):
    """max over channels of max|got-ref| / rms(ref) ; arrays [epochs][ch][t]"""
    rms = np.sqrt((ref.astype(np.float64) ** 2).mean(axis=(0, 2)))
    err = np.abs(got.astype(np.float64) - ref).max(axis=(0, 2))
    return float((err / rms).max())


def test_small_case_vs_reference_golden(golden):
    import eeg_oracle as O
    from eav_b200.ops import PreprocEngine
    g = golden("preproc_small.npz")
    raw = g["raw"]                                     # [6 trials][4 ch][1000]
    n_tr, n_ch, tl = raw.shape
    eng = PreprocEngine(1, n_trials=n_tr, n_chans=n_ch, trial_len=tl)
    keep = np.ones(n_tr, dtype=bool)
    slot = torch.from_numpy(_slots(keep)[None]).cuda()
    taps = O.decimation_taps(5)
    for tag, band in (("b0545", [0.5, 45]), ("b0530", [5, 30])):
        sos = O.butter_sos(band, 100.0)
        ep, dec = eng.run(torch.from_numpy(raw[None]).cuda(), taps, sos, slot, 4 * n_tr, want_dec=True)
        ref_dec = np.transpose(g["dec"], (0, 2, 1)).reshape(n_ch, -1)          # (ch, trials*t)
        d = dec[0].cpu().numpy()
        assert np.abs(d - ref_dec).max() / np.sqrt((ref_dec ** 2).mean()) < 2e-6
        ref = g["filt_" + tag]                                                # (ch, 200, trials)
        ref_ep = np.transpose(ref, (2, 0, 1)).reshape(n_tr, n_ch, 4, 50).transpose(0, 2, 1, 3).reshape(4 * n_tr, n_ch, 50)
        assert _chan_rel_err(ep[0].cpu().numpy(), ref_ep) < TOL, tag


def test_generic_decimation_factor_vs_oracle():
    """down=4 (81 taps) goes through the generic FIR kernel; 3 biquads through another SOS instantiation."""
    import eeg_oracle as O
    from scipy.signal import butter
    from eav_b200.ops import PreprocEngine
    rng = np.random.default_rng(3)
    raw = rng.standard_normal((5, 3, 800)).astype(np.float32) + 2.0
    taps = O.decimation_taps(4)
    sos = np.ascontiguousarray(butter(3, [1, 40], btype="bandpass", fs=125.0, output="sos"))
    eng = PreprocEngine(1, n_trials=5, n_chans=3, trial_len=800, down=4, n_taps=81, n_sections=3, n_sub=4)
    keep = np.array([1, 0, 1, 1, 0], dtype=bool)
    ep, dec = eng.run(torch.from_numpy(raw[None]).cuda(), taps, sos, torch.from_numpy(_slots(keep)[None]).cuda(),
                      4 * 3, want_dec=True)
    ref_dec = O.fir_decimate(raw, taps, 4)
    assert np.abs(dec[0].cpu().numpy() - ref_dec).max() < 2e-6 * np.abs(ref_dec).max()
    filt = O.sosfilt(sos, ref_dec).reshape(3, 5, 4, 50)                       # ch, trial, sub, t
    ref_ep = filt[:, keep].transpose(1, 2, 0, 3).reshape(12, 3, 50)
    assert _chan_rel_err(ep[0].cpu().numpy(), ref_ep) < TOL


@pytest.fixture(scope="module")
def subject_pair():
    import eeg_oracle as O
    raws, labels = zip(*[O.synth_subject(s) for s in (1, 2)])
    return np.stack(raws), labels


def test_dataset_shaped_subjects_vs_oracle_and_reference_digest(golden, subject_pair):
    import eeg_oracle as O
    from eav_b200.ops import PreprocEngine
    raws, labels = subject_pair
    g = golden("preproc_subject1_digest.npz")
    taps, sos = O.decimation_taps(5), O.butter_sos([0.5, 45], 100.0)
    plans = [O.epoch_plan(l) for l in labels]
    slot = np.stack([_slots(p[0]) for p in plans])
    eng = PreprocEngine(2)
    ep = eng.run(torch.from_numpy(raws).cuda(), taps, sos, torch.from_numpy(slot).cuda(), 400).cpu().numpy()
    assert ep.shape == (2, 400, 30, 500)
    # labels / epoch order are host integer logic: bit-exact against the reference
    assert np.array_equal(plans[0][1], g["y"])
    # subject 1 against the reference's own output (strided sample + per-epoch sums + channel RMS)
    rms = g["x_chan_rms"]
    sub = ep[0][::25, ::7, ::20]
    assert (np.abs(sub - g["x_sub"]).max(axis=(0, 2)) / rms[::7]).max() < TOL
    assert np.abs(ep[0].astype(np.float64).sum(axis=(1, 2)) - g["x_epoch_sum"]).max() < 1e-2
    # both subjects against the full CPU oracle
    for s in range(2):
        xo, yo = O.prepare_data(raws[s], labels[s], [0.5, 45])
        assert _chan_rel_err(ep[s], xo) < TOL, s

def test_band_pass_pattern_kernel_equals_general_kernel(subject_pair, monkeypatch):
    """The Butterworth band-pass numerators b0*(1, +-2, 1) / (1, 0, -1) select a kernel instance with the pattern compiled in
    (20 instead of 25 fp64 operations per sample); EAV_SOS_PATTERN=0 forces the general instance.  Same fp64
    products up to the order of two additions: the float32 epochs agree to rounding."""
    import eeg_oracle as O
    from eav_b200.ops import PreprocEngine
    raws, labels = subject_pair
    taps = O.decimation_taps(5)
    slot = torch.from_numpy(np.stack([_slots(O.epoch_plan(l)[0]) for l in labels])).cuda()
    raw = torch.from_numpy(raws).cuda()
    eng = PreprocEngine(2)
    for band in ([0.5, 45], [1, 40]):                 # both take the forgetting-filter path
        sos = O.butter_sos(band, 100.0)
        assert np.array_equal(sos[1:, 0], np.ones(4)) and np.array_equal(np.abs(sos[:, 1] / sos[:, 0]), [2, 2, 0, 2, 2])
        monkeypatch.delenv("EAV_SOS_PATTERN", raising=False)
        a = eng.run(raw, taps, sos, slot, 400).cpu().numpy()
        monkeypatch.setenv("EAV_SOS_PATTERN", "0")
        b = eng.run(raw, taps, sos, slot, 400).cpu().numpy()
        assert _chan_rel_err(a[0], b[0].astype(np.float64)) < 1e-6 and _chan_rel_err(a[1], b[1].astype(np.float64)) < 1e-6


def test_legacy_order_small_vs_oracle():
    """order=1: band-pass at fs_orig over the raw recording, then decimate (CNN_EEG_tf.py:64-75,182-189)."""
    import eeg_oracle as O
    from eav_b200.ops import PreprocEngine
    rng = np.random.default_rng(11)
    n_tr, n_ch, tl = 7, 3, 2000
    raw = rng.standard_normal((n_tr, n_ch, tl)).astype(np.float32)
    raw += (2.0 * np.sin(2 * np.pi * 0.3 * np.arange(n_tr * tl) / 500.0)).reshape(n_tr, 1, tl).astype(np.float32)
    keep = np.array([1, 0, 1, 1, 0, 0, 1], dtype=bool)
    label = np.zeros((10, n_tr)); label[np.where(keep, 1, 0), np.arange(n_tr)] = 1     # class 1 kept, class 0 dropped
    eng = PreprocEngine(1, n_trials=n_tr, n_chans=n_ch, trial_len=tl, order=1)
    sos = O.butter_sos((3, 50), 500.0)
    ep, dec = eng.run(torch.from_numpy(raw[None]).cuda(), O.decimation_taps(5), sos,
                      torch.from_numpy(_slots(keep)[None]).cuda(), 4 * int(keep.sum()), want_dec=True)
    seqs = np.transpose(raw, (1, 0, 2)).reshape(n_ch, n_tr * tl).astype(np.float64)
    filt = O.sosfilt(sos, seqs)
    dec_o = O.fir_decimate(filt.reshape(n_ch, n_tr, tl).transpose(1, 0, 2), O.decimation_taps(5), 5)
    rms = np.sqrt((dec_o ** 2).mean(axis=1))
    assert (np.abs(dec.cpu().numpy()[0] - dec_o).max(axis=1) / rms).max() < TOL
    xo, _ = O.epoch_gather(dec_o, label, tl // 5, 4)
    assert _chan_rel_err(ep.cpu().numpy()[0], xo) < TOL


def test_legacy_order_dataset_shaped_vs_reference_digest(golden, subject_pair):
    """Subject 1 through the drop-in's prepare_data_legacy_device() against the output of the reference's own
    Bandpass()/mysplit() (tests/golden/preproc_legacy_subject1_digest.npz) and against the oracle."""
    import eeg_oracle as O
    from eav_b200.Dataload_eeg import DataLoadEEG
    raws, labels = subject_pair
    g = golden("preproc_legacy_subject1_digest.npz")
    D = DataLoadEEG(subject=1)
    D.set_raw(np.transpose(raws[0], (2, 1, 0)), labels[0])       # (time, ch, trials) as loadmat returns it
    x_dev, y = D.prepare_data_legacy_device(band=(3, 50))
    x = x_dev.cpu().numpy()
    assert x.shape == (400, 30, 500)
    assert np.array_equal(np.eye(5, dtype=np.uint8)[y].T, g["onehot"])        # labels bit-exact, classes 0..4
    rms = g["x_chan_rms"]
    assert (np.abs(x[::25, ::7, ::20] - g["x_sub"]).max(axis=(0, 2)) / rms[::7]).max() < TOL
    assert np.abs(x.astype(np.float64).sum(axis=(1, 2)) - g["x_epoch_sum"]).max() < 1e-2
    xo, yo, _ = O.prepare_data_legacy(raws[0], labels[0], (3, 50))
    assert np.array_equal(y, yo) and _chan_rel_err(x, xo) < TOL
    # the paper's split: 56 train / 24 test epochs per class
    from eav_b200.EAV_datasplit import EAVDataSplit
    tr_x, tr_y, te_x, te_y = EAVDataSplit(x, y).get_split(h_idx=56)
    assert tr_x.shape == (280, 30, 500) and te_x.shape == (120, 30, 500)


def test_legacy_order_rejects_float64():
    from eav_b200.ops import PreprocEngine
    with pytest.raises(RuntimeError):
        PreprocEngine(1, n_trials=2, n_chans=2, trial_len=100, raw_dtype=torch.float64, order=1)


def test_properties_at_full_size(subject_pair):
    """Size-independent properties at the dataset shape: exact homogeneity under power-of-two
    scaling, and independence of a kept epoch from WHICH other trials are kept (the filter is
    continuous over all trials, SURVEY F8, so dropping trials must not change survivors)."""
    import eeg_oracle as O
    from eav_b200.ops import PreprocEngine
    raws, labels = subject_pair
    taps, sos = O.decimation_taps(5), O.butter_sos([5, 30], 100.0)
    keep = O.epoch_plan(labels[0])[0]
    eng = PreprocEngine(1)
    raw = torch.from_numpy(raws[:1]).cuda()
    a = eng.run(raw, taps, sos, torch.from_numpy(_slots(keep)[None]).cuda(), 400).clone()
    b = eng.run(raw * 4.0, taps, sos, torch.from_numpy(_slots(keep)[None]).cuda(), 400).clone()
    assert torch.equal(a * 4.0, b)
    allk = np.ones(200, dtype=bool)
    c = eng.run(raw, taps, sos, torch.from_numpy(_slots(allk)[None]).cuda(), 800)
    kept_idx = np.nonzero(keep)[0]
    sel = (kept_idx[:, None] * 4 + np.arange(4)[None]).reshape(-1)
    assert torch.equal(c[0, torch.from_numpy(sel).cuda()], a[0])
    # a constant record decimates to the same constant away from the record edges (sum h == 1)
    const = torch.full((1, 200, 30, 10000), 3.25, device="cuda")
    _, dec = eng.run(const, taps, sos, torch.from_numpy(_slots(keep)[None]).cuda(), 400, want_dec=True)
    assert (dec[0, :, 20:-20] - 3.25).abs().max().item() < 1e-6


def test_empty_selection_and_errors():
    import eeg_oracle as O
    from eav_b200.ops import PreprocEngine
    eng = PreprocEngine(1, n_trials=4, n_chans=2, trial_len=1000)
    raw = torch.randn(1, 4, 2, 1000, device="cuda")
    slot = torch.full((1, 4), -1, dtype=torch.int32, device="cuda")
    ep = eng.run(raw, O.decimation_taps(5), O.butter_sos([5, 30], 100.0), slot, 0)
    assert ep.shape == (1, 0, 2, 50)
    with pytest.raises(RuntimeError, match="multiple of n_sub"):
        eng.run(raw, O.decimation_taps(5), O.butter_sos([5, 30], 100.0), slot, 3)
    with pytest.raises(RuntimeError):
        eng.run(raw.cpu(), O.decimation_taps(5), O.butter_sos([5, 30], 100.0), slot, 0)
