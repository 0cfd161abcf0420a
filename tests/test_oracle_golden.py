"""Pins the CPU oracle (oracle/) against golden vectors produced by the UNMODIFIED
reference (oracle/gen_golden.py) and against the known-answer constants of
SURVEY.md section 8c.  CPU only."""
import numpy as np
import pytest
import torch

import eeg_oracle as O
import eegnet_oracle as EO
import golden_inputs as GI
from conftest import rel_l2, rel_max


# ---------------------------------------------------------------- known answers
def test_decimation_taps_known_answer():
    h = O.decimation_taps(5)
    assert h.size == 101
    assert h[50] == pytest.approx(0.20013010331628728, abs=1e-16)
    assert np.allclose(h, h[::-1], atol=0, rtol=0)
    assert h.sum() == pytest.approx(1.0, abs=1e-15)
    assert np.abs(h[55]) < 1e-16          # every 5th off-centre tap ~ 0
    from scipy.signal import firwin
    assert np.abs(h - firwin(101, 0.2, window=("kaiser", 5.0))).max() < 1e-17


def test_butter_sos_known_answer():
    sos = O.butter_sos([0.5, 45], 100.0)
    ref = np.array([[0.5697113391879263, 1.1394226783758525, 0.5697113391879263, 1, 1.5207773594869016, 0.5995645436714132],
                    [1, 2, 1, 1, 1.7368413144215746, 0.8263094651760179],
                    [1, 0, -1, 1, -0.24355497400690285, -0.7028117712403572],
                    [1, -2, 1, 1, -1.949393716288919, 0.9503623888786786],
                    [1, -2, 1, 1, -1.9798720600607025, 0.9808504113933361]])
    assert np.abs(sos - ref).max() < 1e-13


def test_sosfilt_c_equals_python_loop():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(300)
    sos = O.butter_sos([5, 30], 100.0)
    assert np.array_equal(O.sosfilt(sos, x[None])[0], O.sosfilt_python(sos, x))


def test_fir_c_equals_numpy_closed_form():
    rng = np.random.default_rng(1)
    raw = rng.standard_normal((3, 2, 250)).astype(np.float32)
    h = O.decimation_taps(5)
    a = O.fir_decimate(raw, h, 5)
    for c in range(2):
        b = O.fir_decimate_numpy(raw[:, c, :].reshape(-1), h, 5)
        assert np.abs(a[c] - b).max() < 1e-14


# ---------------------------------------------------------------- preprocessing vs reference
def test_preproc_small_vs_reference(golden):
    g = golden("preproc_small.npz")
    raw = g["raw"]                                   # [trials][ch][time]
    dec = O.fir_decimate(raw, O.decimation_taps(5), 5)
    ref_dec = g["dec"]                               # (ch, t, trials)
    ref_seq = np.transpose(ref_dec, (0, 2, 1)).reshape(ref_dec.shape[0], -1)
    assert np.abs(dec - ref_seq).max() < 1e-13
    for tag, band in (("b0545", [0.5, 45]), ("b0530", [5, 30])):
        f = O.sosfilt(O.butter_sos(band, 100.0), dec)
        ref = g["filt_" + tag]
        ref_seq = np.transpose(ref, (0, 2, 1)).reshape(ref.shape[0], -1)
        assert np.abs(f - ref_seq).max() < 1e-11


def test_preproc_subject1_digest_vs_reference(golden):
    g = golden("preproc_subject1_digest.npz")
    raw, label = O.synth_subject(1)
    assert np.array_equal(label.astype(np.uint8), g["label"])
    chk = np.array([float(raw.astype(np.float64).sum()), float(np.abs(raw).astype(np.float64).sum())])
    assert np.allclose(chk, g["raw_checksum"], rtol=1e-12), "synthetic generator drifted"
    x, y = O.prepare_data(raw, label, [0.5, 45])
    assert x.shape == (400, 30, 500) and np.array_equal(y, g["y"])
    assert set(np.unique(y)) == {1, 3, 5, 7, 9}
    assert np.abs(x[::25, ::7, ::20] - g["x_sub"]).max() < 1e-11
    assert np.abs(x.sum(axis=(1, 2)) - g["x_epoch_sum"]).max() < 1e-8


def test_legacy_order_digest_vs_reference(golden):
    """Band-pass at 500 Hz BEFORE the decimation (CNN_EEG_tf.py:64-75,180-206): the fixture was produced by the
    reference's own Bandpass()/mysplit() (oracle/gen_golden.py:gen_legacy)."""
    g = golden("preproc_legacy_subject1_digest.npz")
    raw, label = O.synth_subject(1)
    assert np.array_equal(label.astype(np.uint8), g["label"])
    x, y5, onehot = O.prepare_data_legacy(raw, label, (3, 50))
    assert x.shape == (400, 30, 500) and set(np.unique(y5)) == {0, 1, 2, 3, 4}
    assert np.array_equal(onehot.astype(np.uint8), g["onehot"])               # 5-class one-hot, 80 epochs per class
    assert np.abs(x[::25, ::7, ::20] - g["x_sub"]).max() < 1e-11
    assert np.abs(x.sum(axis=(1, 2)) - g["x_epoch_sum"]).max() < 1e-8
    assert np.abs(np.sqrt((x ** 2).mean(axis=(0, 2))) - g["x_chan_rms"]).max() < 1e-11


def test_epoch_plan_vs_reference(golden):
    g = golden("segment_plan.npz")
    keep, y, src_trial, src_sub = O.epoch_plan(g["label"].astype(np.float64))
    assert np.array_equal(y, g["y"])
    assert np.array_equal(src_trial, g["src_trial"])
    assert np.array_equal(src_sub * 500, g["src_t0"])


@pytest.mark.parametrize("tag", ["shipped", "remap"])
def test_split_vs_reference(golden, tag):
    g = golden(f"split_{tag}.npz")
    y = g["y"]
    for h in (40, 56):
        tr, te = O.split_indices(y, h)
        assert np.array_equal(tr, g[f"tr_idx_{h}"]) and np.array_equal(te, g[f"te_idx_{h}"])
        assert np.array_equal(y[tr], g[f"tr_y_{h}"]) and np.array_equal(y[te], g[f"te_y_{h}"])
    if tag == "shipped":
        tr, te = O.split_indices(y, 56)
        assert tr.size == 112 and te.size == 48
    else:
        tr, te = O.split_indices(y, 56)
        assert tr.size == 280 and te.size == 120


# ---------------------------------------------------------------- EEGNet vs reference
def _init(g, variant):
    sd = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init::")}
    return EO.split_state(sd, variant)


@pytest.mark.parametrize("tag", ["b8", "b8_renorm"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_eegnet_tor_fwd_bwd_vs_reference(golden, tag, mode):
    g = golden(f"eegnet_tor_{tag}.npz")
    params, buffers = _init(g, "tor")
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    masks = [torch.from_numpy(g["mask1"]), torch.from_numpy(g["mask2"])] if mode == "train" else None
    p = EO.tor_forward(params, buffers, x, mode == "train", masks=masks)
    loss = EO.loss_fn(p, y)
    loss.backward()
    assert rel_max(p.detach().numpy(), g[f"{mode}::probs"]) < 2e-6
    assert abs(float(loss.detach()) - float(g[f"{mode}::loss"])) < 1e-6
    for k in EO.TOR_PARAMS:
        assert rel_l2(params[k].grad.numpy(), g[f"{mode}::grad::{k}"]) < 2e-5, k
    for k in g.files:
        if k.startswith(f"{mode}::after::"):
            name = k.split("::")[-1]
            mine = (params[name].detach() if name in params else buffers[name]).numpy()
            assert np.allclose(mine, g[k], rtol=1e-5, atol=1e-7), name
    if tag == "b8_renorm":   # the max-norm branch really fired
        assert not np.allclose(g[f"{mode}::after::dense.weight"], g["init::dense.weight"])


def test_dropout_mask_replay_matches_reference_rng(golden):
    """The reference draws mask1 then mask2 from the global CPU RNG."""
    g = golden("eegnet_tor_b8.npz")
    params, buffers = _init(g, "tor")
    torch.manual_seed(100)
    rec = {"record": []}
    p = EO.tor_forward(params, buffers, torch.from_numpy(g["x"]), True, masks=rec)
    assert np.array_equal(rec["record"][0].numpy().astype(np.uint8), g["mask1"])
    assert np.array_equal(rec["record"][1].numpy().astype(np.uint8), g["mask2"])
    assert rel_max(p.detach().numpy(), g["train::probs"]) < 2e-6


def test_adam_trajectory_vs_reference(golden):
    g = golden("eegnet_tor_adam6.npz")
    xs, ys = GI.adam6_inputs()
    assert np.allclose(GI.checksum(xs.numpy(), ys.numpy()), g["input_checksum"], rtol=1e-12)
    params, buffers = _init(g, "tor")
    opt = EO.Adam(params, lr=1e-3)
    losses = []
    for i in range(6):
        masks = [torch.from_numpy(g["masks1"][i]), torch.from_numpy(g["masks2"][i])] if i < 2 else None
        loss, _ = EO.train_step("tor", params, buffers, opt, xs[i], ys[i], i < 2, masks=masks)
        losses.append(float(loss))
    assert np.abs(np.array(losses) - g["losses"]).max() < 2e-5
    for k in EO.TOR_PARAMS:
        # Adam's first steps are sign-like: compare updates loosely, weights tightly
        assert np.abs(params[k].detach().numpy() - g[f"final::{k}"]).max() < 2e-3, k
        assert rel_l2(params[k].detach().numpy(), g[f"final::{k}"]) < 5e-3, k


def test_trainer_uni_loop_vs_reference(golden):
    g = golden("trainer_uni_3ep.npz")
    data = GI.trainer_inputs()
    assert np.allclose(GI.checksum(*data), g["input_checksum"], rtol=1e-12)
    params, buffers = _init(g, "tor")
    torch.manual_seed(77)
    log = EO.trainer_uni_train(params, buffers, data, lr=1e-3, batch_size=16, num_epochs=3)
    ref = g["train_step_loss"]
    assert len(log["step_loss"]) == ref.size == 9
    assert np.abs(np.array(log["step_loss"]) - ref).max() < 5e-5
    assert int(buffers["firstBN.num_batches_tracked"]) == 3          # only epoch 1 trains BN (F5)


def test_trainer_uni_bench_configuration_vs_reference(golden):
    """The benchmark's own configuration (one full subject, 280/120, batch 32 incl. the ragged 24, lr 1e-5, epoch 1
    train mode + epoch 2 eval mode) pins the oracle where bench.py runs."""
    g = golden("trainer_uni_bench_2ep.npz")
    data = GI.bench_subject_inputs()
    assert np.allclose(GI.checksum(*data), g["input_checksum"], rtol=1e-12)
    params, buffers = _init(g, "tor")
    torch.manual_seed(78)
    log = EO.trainer_uni_train(params, buffers, data, lr=1e-5, batch_size=32, num_epochs=2)
    ref = g["train_step_loss"]
    assert len(log["step_loss"]) == ref.size == 18
    assert np.abs(np.array(log["step_loss"]) - ref).max() < 1e-4 * np.abs(ref).max()
    vref = g["val_batch_loss"].reshape(2, 4).mean(1)
    assert np.allclose([v[0] for v in log["val"]], vref, rtol=1e-4)
    assert int(buffers["firstBN.num_batches_tracked"]) == 9          # the 9 steps of epoch 1 only (F5)


@pytest.mark.parametrize("tag", ["default", "eav"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_cnn_eeg_fwd_bwd_vs_reference(golden, tag, mode):
    g = golden(f"cnn_eeg_{tag}.npz")
    params, buffers = _init(g, "cnn")
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    masks = [torch.from_numpy(g["mask1"]), torch.from_numpy(g["mask2"])] if mode == "train" else None
    cfg = dict(dropoutRate=0.25 if tag == "default" else 0.5)
    o = EO.cnn_forward(params, buffers, x, mode == "train", cfg=cfg, masks=masks)
    loss = EO.loss_fn(o, y)
    loss.backward()
    assert rel_max(o.detach().numpy(), g[f"{mode}::logits"]) < 5e-6
    assert abs(float(loss.detach()) - float(g[f"{mode}::loss"])) < 2e-6
    for k in EO.CNN_PARAMS:
        # block1.1 (BN1) feeds BN2 through a linear map with no ELU in between, so in
        # train mode its affine gradient is analytically ~0 and both sides hold rounding
        # noise (|g| ~ 1e-6): gate on an absolute floor there.
        a, b = params[k].grad.numpy(), g[f"{mode}::grad::{k}"]
        assert rel_l2(a, b) < 5e-5 or np.abs(a - b).max() < 5e-6, k


def test_shallowconvnet_oracle_vs_reference(golden):
    """SURVEY 8f.3 groundwork: oracle/shallow_oracle.py against the unmodified Transformer_EEG.ShallowConvNet
    (fixture from oracle/gen_golden.py:gen_shallow): eval forward, train forward/loss/backward with the reference's
    recorded dropout masks, BatchNorm running statistics."""
    import torch
    import shallow_oracle as SO
    g = golden("shallowconvnet_b4.npz")
    sd = {k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("init::")}
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    with torch.no_grad():
        pe = SO.shallow_forward({k: v.clone() for k, v in sd.items()}, x, train=False)
    assert np.abs(pe.numpy() - g["eval::probs"]).max() < 2e-6
    masks = []
    i = 0
    while f"train::mask{i:02d}" in g.files:
        shape = tuple(int(v) for v in g[f"train::mask{i:02d}_shape"])
        bits = np.unpackbits(g[f"train::mask{i:02d}"])[:int(np.prod(shape))].reshape(shape)
        masks.append(torch.from_numpy(bits.astype(np.float32)) * 2.0)          # keep / (1 - 0.5)
        i += 1
    assert len(masks) == 3 * SO.N_LAYERS + 1
    tsd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    pt = SO.shallow_forward(tsd, x, train=True, masks=masks)
    loss = SO.loss_fn(pt, y)
    loss.backward()
    assert np.abs(pt.detach().numpy() - g["train::probs"]).max() < 5e-6
    assert abs(float(loss) - float(g["train::loss"])) < 1e-6
    for k in g.files:
        if k.startswith("train::grad::"):
            ref = g[k]
            got = tsd[k[len("train::grad::"):]].grad.numpy()
            denom = max(np.linalg.norm(ref), 1e-8)
            assert np.linalg.norm(got - ref) / denom < 2e-4, k
    assert np.allclose(tsd["bn.running_mean"].numpy(), g["train::bn_running_mean"], rtol=1e-5, atol=1e-6)
    assert np.allclose(tsd["bn.running_var"].numpy(), g["train::bn_running_var"], rtol=1e-5, atol=1e-6)
